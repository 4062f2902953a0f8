/* b2d.h — C ABI of libb2d.so: the sm_100a DiT-training-step kernels behind finetrainers' LTX hot path.
 *
 * The reference (a-r-r-o-w/finetrainers @ f476c37) is pure Python and has NO FFI; every entry point below replaces a
 * span of PyTorch/diffusers/peft calls on the hot path.  The citation after each declaration names that span
 * (paths relative to /root/reference; "diffusers:"/"peft:" = the un-vendored dependency the reference delegates to).
 *
 * Conventions: plain pointers + sizes, no torch types, no hidden allocation, no implicit synchronisation.  All device
 * pointers are 16-byte aligned, activations/weights bf16 row-major, statistics/gradients fp32.  The last argument is
 * the CUDA stream (cudaStream_t passed as void*).  Return 0 on success, negative b2d_status on error
 * (b2d_last_error() gives a thread-local message).  Callable from any host thread (autograd's backward thread too).
 */
#ifndef B2D_H
#define B2D_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B2D_OK = 0,
  B2D_ERR_SHAPE = -1,   /* unsupported / inconsistent dims */
  B2D_ERR_ALIGN = -2,   /* pointer or leading dimension not 16-byte aligned */
  B2D_ERR_ARCH = -3,    /* device is not sm_100 */
  B2D_ERR_CUDA = -4,    /* CUDA runtime/driver error (see b2d_last_error) */
  B2D_ERR_ARG = -5
} b2d_status;

int b2d_version(void);                 /* ABI version (this header = 1) */
const char* b2d_last_error(void);      /* thread-local, never NULL */
int b2d_device_check(void);            /* B2D_OK iff current device is compute capability 10.x */

/* ---------------------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores (TMA -> 128B-swizzled smem -> tcgen05.mma -> TMEM -> fused epilogue).
 *   C[M,N] = epilogue( alpha * ( opA(A)[M,K] * opB(B)[N,K]^T  +  A2[M,K2] * B2[N,K2]^T ) )
 * a_mn_major = 0: A is row-major [M, K] (lda);  1: A is given as its transpose, row-major [K, M] (lda)
 * b_mn_major = 0: B is row-major [N, K] (ldb) (an nn.Linear weight); 1: row-major [K, N] (ldb)
 * The optional (A2,B2) pair extends the contraction by K2 (LoRA low-rank update fused into the same accumulator):
 *   a2_group_n > 0 : the A2 column offset for output tile column n0 is (n0 / a2_group_n) * K2 (per-adapter slices of a
 *                    packed u = [u_q|u_k|u_v]); 0 : offset 0.
 * splits > 1 splits the main contraction over CTAs; only valid with the fp32 atomic epilogues.
 * batch > 1 repeats the problem with per-batch element offsets (a_boff, b_boff, c_boff... applied as coordinates).
 * Replaces: every nn.Linear on the path (diffusers: LTXVideoTransformerBlock / Attention / FeedForward;
 *   finetrainers/patches/models/ltx_video/patch.py:82-85,118-123), peft: lora.Linear.forward, and their autograd
 *   backward (dX; LoRA dA/dB).
 * ------------------------------------------------------------------------------------------------------------- */
typedef enum {
  B2D_EPI_STORE = 0,        /* out(bf16) = alpha*acc + bias */
  B2D_EPI_GELU = 1,         /* pre = acc + bias; out2 = pre (optional); out = gelu_tanh(pre) */
  B2D_EPI_SILU = 2,         /* pre = acc + bias; out2 = pre (optional); out = silu(pre) */
  B2D_EPI_GATE_RES = 3,     /* out = res + gate[b,col] * (acc + bias); gate = gate_table[col] + gate_temb[b,col] or 1;
                               optional out2 = out * (gate2_table[col] + gate2_temb[b,col]) */
  B2D_EPI_MUL_DGELU = 4,    /* out = acc * gelu_tanh'(aux) */
  B2D_EPI_F32_ATOMIC = 5,   /* out_f32[row, col]  += alpha*acc   (split-K) */
  B2D_EPI_F32_ATOMIC_T = 6, /* out_f32[col, row]  += alpha*acc   (transposed accumulate) */
  B2D_EPI_F32_STORE = 7     /* out_f32[row, col]   = alpha*acc + bias */
} b2d_epilogue;

typedef struct {
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  const void* A2; int64_t lda2;
  const void* B2; int64_t ldb2;
  int32_t M, N, K, K2;
  int32_t a_mn_major, b_mn_major;
  int32_t a2_group_n;
  int32_t splits, batch;
  int64_t a_boff_row, a_boff_col, b_boff_row, b_boff_col, c_boff;  /* per-batch offsets (elements / rows / cols) */
  int32_t epi;
  float alpha;               /* multiplies the accumulator; set it explicitly (1.0f for a plain product; 0 yields zeros) */
  void* out; int64_t ldc;
  void* out2; int64_t ldc2;
  const void* bias;                 /* bf16 [N] or NULL */
  const void* res; int64_t ldres;   /* bf16 [M, N] */
  const void* aux; int64_t ldaux;   /* bf16 [M, N] */
  const void* gate_table;           /* bf16 [N]          (row of scale_shift_table) or NULL */
  const void* gate_temb;            /* bf16 [nb, temb_stride] (already offset to the gate row) or NULL */
  const void* gate2_table;
  const void* gate2_temb;
  int64_t temb_stride;
  int32_t rows_per_sample;          /* b = row / rows_per_sample */
  int32_t block_n;                  /* 0 = auto; else 64/128/160/192/256 */
  int32_t max_ctas;                 /* 0 = #SMs */
  /* batch offsets of the extension operands and of the bias (batch z reads A2 rows + z*a2_boff_row, B2 rows
   * + z*b2_boff_row, bias + z*bias_boff elements): one launch covers the same projection of several DiT blocks */
  int64_t a2_boff_row, b2_boff_row, bias_boff;
  /* tile scheduling across CTAs: 0 = auto (cost model), 1 = one CTA per 128 x block_n tile, 2 = CTA pairs sharing a
   * 256 x block_n tile (tcgen05 cta_group::2; needs K-major A, no split-K, block_n in 128/160/192/256, M > 128) */
  int32_t cta_pair;
} b2d_gemm_desc;

int b2d_gemm(const b2d_gemm_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused RMSNorm / LayerNorm (no affine) + AdaLN modulate.   y = norm(x) * (1 + scale[b]) + shift[b]
 *   scale[b,c] = table[scale_row, c] + temb[b, scale_row*D + c]  (likewise shift); layer_norm=1 subtracts the mean.
 * Replaces: diffusers LTXVideoTransformerBlock norm1/norm2 + ada_values (and patch.py:113-120 norm_out + modulate);
 *   RMSNorm numerics finetrainers/patches/dependencies/diffusers/rms_norm.py:17-30.
 * bwd: dx_accum += d norm/dx ( dy * (1+scale) )   (adds into the residual-stream gradient; optional second output
 *   dx_scaled = dx_accum * gate2[b] for the next GEMM's A operand)
 * Row kernels: one row per 256-thread CTA; D must be a multiple of 8 and <= 8192.
 * ------------------------------------------------------------------------------------------------------------- */
int b2d_norm_modulate_fwd(const void* x, void* y, const void* shift_tab, const void* shift_emb, const void* scale_tab,
                          const void* scale_emb, int64_t emb_stride, int32_t rows, int32_t D, int32_t rows_per_sample,
                          float eps, int32_t layer_norm, void* stream);
/* dx_out = (accumulate ? dx_accum_in : 0) + dnorm(dy * (1 + scale)); optional out2 = dx_out * (gate2_tab + gate2_emb[b]) */
int b2d_norm_modulate_bwd(const void* dy, const void* x, const void* dx_in, void* dx_out, const void* scale_tab,
                          const void* scale_emb, const void* gate2_tab, const void* gate2_emb, void* out2,
                          int64_t emb_stride, int32_t rows, int32_t D, int32_t rows_per_sample, float eps,
                          int32_t layer_norm, void* stream);

/* out = x * (tab[c] + emb[b, c]) per column (gate application on the gradient path). */
int b2d_colscale(const void* x, void* out, const void* tab, const void* emb, int64_t emb_stride, int32_t rows,
                 int32_t D, int32_t rows_per_sample, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * q/k RMSNorm-across-heads (affine) + 3-D RoPE + head split.
 *   src [rows, ld] bf16 (q, k, v at column offsets) -> q',k',v' in [B, H, S, 64].
 * rope cos/sin: fp32 [S, D/2], one value per rotary pair (NULL = no RoPE: cross attention).
 * Replaces: diffusers LTXVideoAttentionProcessor2_0 (norm_q/norm_k, apply_rotary_emb patch.py:23-33, unflatten+transpose).
 * ------------------------------------------------------------------------------------------------------------- */
int b2d_qknorm_rope_fwd(const void* src, int64_t ld, int64_t col_off, const void* weight, const void* cos,
                        const void* sin, void* dst, int32_t B, int32_t S, int32_t H, int32_t norm, float eps,
                        void* stream);
int b2d_qknorm_rope_bwd(const void* dsrc_heads, const void* x, int64_t ld, int64_t col_off, const void* weight,
                        const void* cos, const void* sin, void* dx, int64_t ld_dx, int64_t dx_col_off, int32_t B,
                        int32_t S, int32_t H, int32_t norm, float eps, void* stream);
/* Same, for nseg (1..3) consecutive D-wide column segments of one packed row in ONE launch (q|k|v of the fused QKV
 * projection; k|v of cross attention): segment i lives at col_off + i*D, is RMS-normed iff w_i != NULL, rotated iff bit i of
 * rope_mask is set, and is written head-split to dst_i.  The (cos, sin) row is read once for all segments.  The backward
 * reads the head-split upstream gradients dy_i and writes dx[row, dx_col_off + i*D + c].
 * rows_per_w > 0: the rows are several DiT blocks stacked (B = blocks * batch); row r then uses the norm weights
 * w_i + (r / rows_per_w) * w_stride (elements) - the text-side k|v of all blocks in one launch. */
int b2d_qkv_norm_rope_fwd(const void* src, int64_t ld, int64_t col_off, int32_t nseg, const void* w0, const void* w1,
                          const void* w2, int32_t rope_mask, const void* cos, const void* sin, void* dst0, void* dst1,
                          void* dst2, int32_t B, int32_t S, int32_t H, float eps, int32_t rows_per_w, int64_t w_stride,
                          void* stream);
int b2d_qkv_norm_rope_bwd(const void* dy0, const void* dy1, const void* dy2, const void* x, int64_t ld, int64_t col_off,
                          int32_t nseg, const void* w0, const void* w1, const void* w2, int32_t rope_mask, const void* cos,
                          const void* sin, void* dx, int64_t ld_dx, int64_t dx_col_off, int32_t B, int32_t S, int32_t H,
                          float eps, int32_t rows_per_w, int64_t w_stride, void* stream);

/* RoPE table (diffusers LTXVideoRotaryPosEmbed.forward, called at patch.py:52): fp32 cos,sin [F*H*W, D/2]
 * (the reference's repeat_interleave(2) duplicates are not stored). */
int b2d_rope_table(float* cos, float* sin, int32_t F, int32_t H, int32_t W, int32_t D, float sf, float sh, float sw,
                   void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Attention, d_head = 64, non-causal, optional additive key bias [B, Sk] (fp32; the -10000 mask bias of patch.py:55-57).
 *   q [B,H,Sq,64], k,v [B,H,Sk,64] bf16 -> out [B,Sq,H*64] bf16 (token-major, feeds to_out directly), lse [B,H,Sq] fp32.
 * Replaces: F.scaled_dot_product_attention == finetrainers/models/attention_dispatch.py:405-447 -> _native_attention
 *   :938-962, and its backward.
 * ------------------------------------------------------------------------------------------------------------- */
int b2d_attn_fwd(const void* q, const void* k, const void* v, const float* key_bias, void* out, float* lse, int32_t B,
                 int32_t H, int32_t Sq, int32_t Sk, float scale, void* stream);
/* dout [B,Sq,H*64] bf16; out as produced by fwd; dq,dk,dv [B,H,S,64] bf16; workspace delta_ws:
 * 2*B*H*Sq floats, plus 2*B*H*Sk*64 + B*H floats when Sk <= 512 (fp32 dK/dV accumulators and per-head arrival
 * counters of the cross-attention paths; the library zeroes what it uses, the caller only provides the space). */
int b2d_attn_bwd(const void* q, const void* k, const void* v, const float* key_bias, const void* out, const void* dout,
                 const float* lse, float* delta_ws, void* dq, void* dk, void* dv, int32_t B, int32_t H, int32_t Sq,
                 int32_t Sk, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Step prologue / epilogue.
 * prep: normalise latents, x_t = (1-sigma) x0 + sigma n (first latent frame may use sigma_ff[b]), pack [B,C,F,H,W] ->
 *   [B, F*H*W, C]; target = n - x0 (packed).   finetrainers/models/ltx_video/base_specification.py:285-322,343,427-459;
 *   finetrainers/functional/diffusion.py:4-11.
 * loss: loss = mean_b( mean_{s,c}( w[b] * (pred - target)^2 ) ) * loss_scale  (fp32);  dpred = dloss/dpred (bf16).
 *   finetrainers/trainer/sft_trainer/trainer.py:463-481.
 * ------------------------------------------------------------------------------------------------------------- */
int b2d_prep_noise_pack(const void* latents, const void* noise, const float* mean, const float* std,
                        const float* sigma, const float* sigma_ff, void* x_t, void* target, int32_t B, int32_t C,
                        int32_t F, int32_t HW, void* stream);
int b2d_loss_mse(const void* pred, const void* target, const float* weight, float loss_scale, float* loss_out,
                 void* dpred, float* partial_ws, int32_t B, int64_t per_sample, void* stream);

/* sinusoidal timestep features (diffusers Timesteps(256, flip_sin_to_cos=True)): out bf16 [n, 256] = [cos | sin]. */
int b2d_timestep_sinusoid(const float* t, void* out, int32_t n, void* stream);

/* fp32 -> bf16 cast with scale (LoRA operand refresh each step). */
int b2d_cast_f32_bf16(const float* src, void* dst, int64_t n, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Flat-buffer optimiser path ("next" row: clip + AdamW; finetrainers/utils/torch.py:99-161, optimizer.py:117-125).
 * ------------------------------------------------------------------------------------------------------------- */
int b2d_sumsq(const float* x, int64_t n, float* out_sumsq /* += */, float* partial_ws, void* stream);
/* p, g, m, v: 16-byte aligned (any n; four elements per thread as 128-bit accesses); g is zeroed (fused zero_grad). */
int b2d_adamw_clip(float* p, float* g, float* m, float* v, int64_t n, const float* sumsq, float max_norm, float lr,
                   float beta1, float beta2, float eps, float wd, int32_t step, float grad_div, void* stream);

#ifdef __cplusplus
}
#endif
#endif
