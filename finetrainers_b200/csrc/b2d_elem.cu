// b2d_elem.cu — the HBM-bound satellites of the DiT step: fused norm+AdaLN modulate (fwd/bwd), q/k RMSNorm + RoPE +
// head split (fwd/bwd), RoPE table, noise/pack prologue, MSE loss + dpred, sinusoid, casts, flat clip + AdamW.
// All: 128-bit coalesced global access, fp32 math, warp-shuffle reductions; one row per CTA of 256 threads.
#include "b2d_internal.h"
#include "b2d_ptx.cuh"

namespace b2d {

constexpr int ROW_THREADS = 256;
constexpr int MAX_CHUNKS = 4;  // D <= 8 * 256 * 4 = 8192

__device__ __forceinline__ uint4 ldg16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stg16(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ void unpack8(uint4 u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// block-wide sum of up to two values; all threads get the result
__device__ __forceinline__ float2 block_sum2(float a, float b) {
    __shared__ float sa[ROW_THREADS / 32], sb[ROW_THREADS / 32];
    __shared__ float ra, rb;
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();  // protect sa/sb/ra/rb reuse across consecutive calls
    if (l == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        float x = l < ROW_THREADS / 32 ? sa[l] : 0.f;
        float y = l < ROW_THREADS / 32 ? sb[l] : 0.f;
        x = warp_sum(x);
        y = warp_sum(y);
        if (l == 0) { ra = x; rb = y; }
    }
    __syncthreads();
    return make_float2(ra, rb);
}

// ------------------------------------------------------------------------------------------------
// norm + modulate
// ------------------------------------------------------------------------------------------------
// Every global operand of a row is requested BEFORE the first block reduction: a row's critical path is then one memory
// round trip + the reductions instead of two or three dependent round trips (these kernels have ~16 KB in flight per
// CTA and 8 CTAs per SM, so the dependent-latency chain, not bandwidth, was what bounded them).
template <int NCH>
__global__ void __launch_bounds__(ROW_THREADS) norm_modulate_fwd_kernel(
    const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ shift_tab,
    const __nv_bfloat16* __restrict__ shift_emb, const __nv_bfloat16* __restrict__ scale_tab,
    const __nv_bfloat16* __restrict__ scale_emb, long long emb_stride, int D, int rows_per_sample, float eps,
    int layer_norm) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x;
    const int b = row / rows_per_sample;
    const __nv_bfloat16* xr = x + (long long)row * D;
    uint4 xq[NCH], q_sht[NCH], q_she[NCH], q_sct[NCH], q_sce[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            xq[c] = ldg16(xr + col);
            q_sht[c] = ldg16(shift_tab + col);
            q_she[c] = ldg16(shift_emb + (long long)b * emb_stride + col);
            q_sct[c] = ldg16(scale_tab + col);
            q_sce[c] = ldg16(scale_emb + (long long)b * emb_stride + col);
        }
    }
    float v[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            unpack8(xq[c], v[c]);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s1 += v[c][e]; s2 += v[c][e] * v[c][e]; }
        }
    }
    float2 tot = block_sum2(s1, s2);
    float mean = 0.f, rstd;
    if (layer_norm) {
        mean = tot.x / D;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = (c * ROW_THREADS + threadIdx.x) * 8;
            if (col < D) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { float d = v[c][e] - mean; var += d * d; }
            }
        }
        float2 t2 = block_sum2(var, 0.f);
        rstd = rsqrtf(t2.x / D + eps);
    } else {
        rstd = rsqrtf(tot.y / D + eps);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            float sh[8], sc[8], t[8];
            unpack8(q_sht[c], sh);
            unpack8(q_she[c], t);
#pragma unroll
            for (int e = 0; e < 8; ++e) sh[e] += t[e];
            unpack8(q_sct[c], sc);
            unpack8(q_sce[c], t);
#pragma unroll
            for (int e = 0; e < 8; ++e) sc[e] += t[e];
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (v[c][e] - mean) * rstd * (1.f + sc[e]) + sh[e];
            stg16(y + (long long)row * D + col, pack8(o));
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(ROW_THREADS) norm_modulate_bwd_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dx_in,
    __nv_bfloat16* __restrict__ dx_out, const __nv_bfloat16* __restrict__ scale_tab,
    const __nv_bfloat16* __restrict__ scale_emb, const __nv_bfloat16* __restrict__ gate2_tab,
    const __nv_bfloat16* __restrict__ gate2_emb, __nv_bfloat16* __restrict__ out2, long long emb_stride, int D,
    int rows_per_sample, float eps, int layer_norm) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x;
    const int b = row / rows_per_sample;
    const long long ro = (long long)row * D;
    uint4 q_x[NCH], q_dy[NCH], q_sct[NCH], q_sce[NCH], q_in[NCH], q_gt[NCH], q_ge[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            q_x[c] = ldg16(x + ro + col);
            q_dy[c] = ldg16(dy + ro + col);
            q_sct[c] = ldg16(scale_tab + col);
            q_sce[c] = ldg16(scale_emb + (long long)b * emb_stride + col);
            q_in[c] = dx_in != nullptr ? ldg16(dx_in + ro + col) : make_uint4(0, 0, 0, 0);
            if (out2 != nullptr) {
                q_gt[c] = ldg16(gate2_tab + col);
                q_ge[c] = ldg16(gate2_emb + (long long)b * emb_stride + col);
            }
        }
    }
    float xv[NCH][8], g[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            unpack8(q_x[c], xv[c]);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s1 += xv[c][e]; s2 += xv[c][e] * xv[c][e]; }
        }
    }
    float2 tot = block_sum2(s1, s2);
    float mean = 0.f, rstd;
    if (layer_norm) {
        mean = tot.x / D;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = (c * ROW_THREADS + threadIdx.x) * 8;
            if (col < D) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { float d = xv[c][e] - mean; var += d * d; }
            }
        }
        rstd = rsqrtf(block_sum2(var, 0.f).x / D + eps);
    } else {
        rstd = rsqrtf(tot.y / D + eps);
    }
    // g = dy * (1 + scale);  xhat = (x - mean) * rstd;  a = sum(g), c = sum(g * xhat)
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            float sc[8], t[8], d[8];
            unpack8(q_sct[c], sc);
            unpack8(q_sce[c], t);
            unpack8(q_dy[c], d);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                g[c][e] = d[e] * (1.f + sc[e] + t[e]);
                xv[c][e] = (xv[c][e] - mean) * rstd;
                sg += g[c][e];
                sgx += g[c][e] * xv[c][e];
            }
        }
    }
    float2 t2 = block_sum2(sg, sgx);
    const float mg = layer_norm ? t2.x / D : 0.f;
    const float mgx = t2.y / D;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            float o[8];
            unpack8(q_in[c], o);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += rstd * (g[c][e] - mg - xv[c][e] * mgx);
            uint4 packed = pack8(o);
            stg16(dx_out + ro + col, packed);
            if (out2 != nullptr) {
                float r[8], gt[8], ge[8];
                unpack8(packed, r);  // the rounded value is what downstream sees
                unpack8(q_gt[c], gt);
                unpack8(q_ge[c], ge);
#pragma unroll
                for (int e = 0; e < 8; ++e) r[e] *= (gt[e] + ge[e]);
                stg16(out2 + ro + col, pack8(r));
            }
        }
    }
}

__global__ void colscale_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                const __nv_bfloat16* __restrict__ tab, const __nv_bfloat16* __restrict__ emb,
                                long long emb_stride, long long total8, int D, int rows_per_sample) {
    griddep_launch_dependents();
    griddep_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total8) return;
    long long e0 = i * 8;
    int row = (int)(e0 / D);
    int col = (int)(e0 % D);
    int b = row / rows_per_sample;
    float v[8], t[8], g[8];
    unpack8(ldg16(x + e0), v);
    unpack8(ldg16(tab + col), t);
    unpack8(ldg16(emb + (long long)b * emb_stride + col), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= (t[e] + g[e]);
    stg16(out + e0, pack8(v));
}

// ------------------------------------------------------------------------------------------------
// q/k RMSNorm (affine, across all heads) + RoPE + head split for up to three column segments of one packed row
// (q | k | v of the fused QKV projection, or k | v of cross attention):  src[row, col_off + i*D + c] -> dst_i[b, h, s, d].
// One CTA handles all segments of its row: the (cos, sin) row is read once for q AND k, every global operand is
// requested before the first reduction, and the two RMS statistics share one block reduction.
// ------------------------------------------------------------------------------------------------
struct QkvSegArgs {
    const __nv_bfloat16* w[3];   // RMSNorm weight of segment i, or null: no norm
    __nv_bfloat16* dst[3];       // fwd: head-split outputs;  bwd: head-split upstream gradients (read)
    int nseg;
    int rope_mask;               // bit i: segment i is rotated
    int rows_per_w;              // > 0: rows are stacked DiT blocks, row r uses weights w[i] + (r / rows_per_w) * w_stride
    long long w_stride;
};

template <int NCH>
__global__ void __launch_bounds__(ROW_THREADS) qkv_norm_rope_fwd_kernel(
    const __nv_bfloat16* __restrict__ src, long long ld, long long col_off, const QkvSegArgs a,
    const float* __restrict__ cosT, const float* __restrict__ sinT, int S, int H, float eps) {
    griddep_launch_dependents();
    griddep_wait();
    const int D = H * 64;
    const int row = blockIdx.x;
    const int b = row / S, s = row % S;
    const __nv_bfloat16* xr = src + (long long)row * ld + col_off;
    const long long woff = a.rows_per_w > 0 ? (long long)(row / a.rows_per_w) * a.w_stride : 0;
    uint4 xq[3][NCH], wq[3][NCH];
    float4 c4[NCH], s4[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (i < a.nseg) {
                    xq[i][c] = ldg16(xr + (long long)i * D + col);
                    if (a.w[i] != nullptr) wq[i][c] = ldg16(a.w[i] + woff + col);
                }
            }
            if (a.rope_mask) {
                // tables hold one (cos, sin) per rotary PAIR: [S, D/2] fp32 (the reference's repeat_interleave(2) is implicit)
                c4[c] = *reinterpret_cast<const float4*>(cosT + ((long long)s * D + col) / 2);
                s4[c] = *reinterpret_cast<const float4*>(sinT + ((long long)s * D + col) / 2);
            }
        }
    }
    float ss[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < a.nseg && a.w[i] != nullptr) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = (c * ROW_THREADS + threadIdx.x) * 8;
                if (col < D) {
                    float v[8];
                    unpack8(xq[i][c], v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) ss[i] += v[e] * v[e];
                }
            }
        }
    }
    float rstd[3] = {1.f, 1.f, 1.f};
    if (a.w[0] != nullptr || (a.nseg > 1 && a.w[1] != nullptr)) {
        const float2 t = block_sum2(ss[0], ss[1]);
        rstd[0] = rsqrtf(t.x / D + eps);
        rstd[1] = rsqrtf(t.y / D + eps);
    }
    if (a.nseg > 2 && a.w[2] != nullptr) rstd[2] = rsqrtf(block_sum2(ss[2], 0.f).x / D + eps);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < a.nseg) {
            const bool norm = a.w[i] != nullptr;
            const bool rope = (a.rope_mask >> i) & 1;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = (c * ROW_THREADS + threadIdx.x) * 8;
                if (col < D) {
                    float n[8];
                    unpack8(xq[i][c], n);
                    if (norm) {
                        float w[8];
                        unpack8(wq[i][c], w);
#pragma unroll
                        for (int e = 0; e < 8; ++e) n[e] = n[e] * rstd[i] * w[e];
                    }
                    float o[8];
                    if (rope) {
                        const float cs[4] = {c4[c].x, c4[c].y, c4[c].z, c4[c].w};
                        const float sn[4] = {s4[c].x, s4[c].y, s4[c].z, s4[c].w};
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            o[e] = n[e] * cs[e >> 1] - n[e + 1] * sn[e >> 1];
                            o[e + 1] = n[e + 1] * cs[e >> 1] + n[e] * sn[e >> 1];
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = n[e];
                    }
                    const int h = col >> 6, d = col & 63;
                    stg16(a.dst[i] + (((long long)b * H + h) * S + s) * 64 + d, pack8(o));
                }
            }
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(ROW_THREADS) qkv_norm_rope_bwd_kernel(
    const __nv_bfloat16* __restrict__ x, long long ld, long long col_off, const QkvSegArgs a,
    const float* __restrict__ cosT, const float* __restrict__ sinT, __nv_bfloat16* __restrict__ dx, long long ld_dx,
    long long dx_col_off, int S, int H, float eps) {
    griddep_launch_dependents();
    griddep_wait();
    const int D = H * 64;
    const int row = blockIdx.x;
    const int b = row / S, s = row % S;
    const __nv_bfloat16* xr = x + (long long)row * ld + col_off;
    const long long woff = a.rows_per_w > 0 ? (long long)(row / a.rows_per_w) * a.w_stride : 0;
    uint4 xq[3][NCH], wq[3][NCH], dq[3][NCH];
    float4 c4[NCH], s4[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * ROW_THREADS + threadIdx.x) * 8;
        if (col < D) {
            const int h = col >> 6, d = col & 63;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (i < a.nseg) {
                    dq[i][c] = ldg16(a.dst[i] + (((long long)b * H + h) * S + s) * 64 + d);
                    if (a.w[i] != nullptr) {
                        xq[i][c] = ldg16(xr + (long long)i * D + col);
                        wq[i][c] = ldg16(a.w[i] + woff + col);
                    }
                }
            }
            if (a.rope_mask) {
                c4[c] = *reinterpret_cast<const float4*>(cosT + ((long long)s * D + col) / 2);
                s4[c] = *reinterpret_cast<const float4*>(sinT + ((long long)s * D + col) / 2);
            }
        }
    }
    float ss[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < a.nseg && a.w[i] != nullptr) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = (c * ROW_THREADS + threadIdx.x) * 8;
                if (col < D) {
                    float v[8];
                    unpack8(xq[i][c], v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) ss[i] += v[e] * v[e];
                }
            }
        }
    }
    float rstd[3] = {1.f, 1.f, 1.f};
    const bool n01 = a.w[0] != nullptr || (a.nseg > 1 && a.w[1] != nullptr);
    const bool n2 = a.nseg > 2 && a.w[2] != nullptr;
    if (n01) {
        const float2 t = block_sum2(ss[0], ss[1]);
        rstd[0] = rsqrtf(t.x / D + eps);
        rstd[1] = rsqrtf(t.y / D + eps);
    }
    if (n2) rstd[2] = rsqrtf(block_sum2(ss[2], 0.f).x / D + eps);
    // g = rope^T(dy) * w ; xhat = x * rstd ; dx = rstd * (g - xhat * mean(g * xhat))
    float g[3][NCH][8];
    float sgx[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < a.nseg) {
            const bool norm = a.w[i] != nullptr;
            const bool rope = (a.rope_mask >> i) & 1;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = (c * ROW_THREADS + threadIdx.x) * 8;
                if (col < D) {
                    float dy[8];
                    unpack8(dq[i][c], dy);
                    if (rope) {
                        const float cs[4] = {c4[c].x, c4[c].y, c4[c].z, c4[c].w};
                        const float sn[4] = {s4[c].x, s4[c].y, s4[c].z, s4[c].w};
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            g[i][c][e] = dy[e] * cs[e >> 1] + dy[e + 1] * sn[e >> 1];
                            g[i][c][e + 1] = dy[e + 1] * cs[e >> 1] - dy[e] * sn[e >> 1];
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) g[i][c][e] = dy[e];
                    }
                    if (norm) {
                        float w[8], xv[8];
                        unpack8(wq[i][c], w);
                        unpack8(xq[i][c], xv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            g[i][c][e] *= w[e];
                            sgx[i] += g[i][c][e] * (xv[e] * rstd[i]);
                        }
                    }
                }
            }
        }
    }
    float mgx[3] = {0.f, 0.f, 0.f};
    if (n01) {
        const float2 t = block_sum2(sgx[0], sgx[1]);
        mgx[0] = t.x / D;
        mgx[1] = t.y / D;
    }
    if (n2) mgx[2] = block_sum2(sgx[2], 0.f).x / D;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < a.nseg) {
            const bool norm = a.w[i] != nullptr;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = (c * ROW_THREADS + threadIdx.x) * 8;
                if (col < D) {
                    float o[8];
                    if (norm) {
                        float xv[8];
                        unpack8(xq[i][c], xv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = rstd[i] * (g[i][c][e] - (xv[e] * rstd[i]) * mgx[i]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = g[i][c][e];
                    }
                    stg16(dx + (long long)row * ld_dx + dx_col_off + (long long)i * D + col, pack8(o));
                }
            }
        }
    }
}

// RoPE table: diffusers LTXVideoRotaryPosEmbed, fp32, one (cos, sin) per rotary pair: [S, D/2].  One thread per (s, pair).
__global__ void rope_table_kernel(float* __restrict__ cosT, float* __restrict__ sinT, int F, int H, int W, int D,
                                  float sf, float sh, float sw) {
    const int nf = D / 6;
    const int pad = D % 6;
    const long long S = (long long)F * H * W;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = D / 2;
    if (idx >= S * pairs) return;
    const int s = (int)(idx / pairs);
    const int pr = (int)(idx % pairs);
    const int col = pr * 2;
    float c = 1.f, sn = 0.f;
    if (col >= pad) {
        const int j = (col - pad) / 2;
        const int fi = j / 3, axis = j % 3;
        const int w = s % W, h = (s / W) % H, f = s / (W * H);
        float g = axis == 0 ? (float)f * sf : (axis == 1 ? (float)h * sh : (float)w * sw);
        // torch.linspace(0, 1, nf) (symmetric evaluation), theta ** x, * pi/2, * (2g - 1): same op order, fp32
        const float step = 1.0f / (float)(nf - 1);
        float lin = (fi < nf / 2) ? (float)fi * step : 1.0f - (float)(nf - 1 - fi) * step;
        float fr = powf(10000.0f, lin);
        fr = fr * 1.5707963267948966f;
        float ang = fr * (g * 2.0f - 1.0f);
        c = cosf(ang);
        sn = sinf(ang);
    }
    cosT[(long long)s * pairs + pr] = c;
    sinT[(long long)s * pairs + pr] = sn;
}

// ------------------------------------------------------------------------------------------------
// step prologue: normalise + flow-match x_t + pack [B,C,F,HW] -> [B, F*HW, C]; target = n - x0.
// Rounding points mirror the reference's bf16 tensors (base_specification.py:295-322): x0 rounded to bf16, x_t computed
// in fp32 from (bf16 x0, bf16 n, fp32 sigma) then rounded, target = bf16(n - x0).
// ------------------------------------------------------------------------------------------------
__global__ void prep_noise_pack_kernel(const __nv_bfloat16* __restrict__ lat, const __nv_bfloat16* __restrict__ noise,
                                       const float* __restrict__ mean, const float* __restrict__ stdv,
                                       const float* __restrict__ sigma, const float* __restrict__ sigma_ff,
                                       __nv_bfloat16* __restrict__ x_t, __nv_bfloat16* __restrict__ target, int B,
                                       int C, int F, int HW) {
    const long long S = (long long)F * HW;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * S * C) return;
    const int c = (int)(idx % C);
    const long long s = (idx / C) % S;
    const int b = (int)(idx / (C * S));
    const long long src = ((long long)b * C + c) * S + s;
    float x = __bfloat162float(lat[src]);
    float x0f = __fdiv_rn(__fmul_rn(__fsub_rn(x, mean[b * C + c]), 1.0f), stdv[b * C + c]);
    float x0 = __bfloat162float(__float2bfloat16_rn(x0f));
    float n = __bfloat162float(noise[src]);
    float sg = sigma[b];
    if (sigma_ff != nullptr && s < HW) sg = sigma_ff[b];
    float xt = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, sg), x0), __fmul_rn(sg, n));
    x_t[idx] = __float2bfloat16_rn(xt);
    target[idx] = __float2bfloat16_rn(__fsub_rn(n, x0));
}

// ------------------------------------------------------------------------------------------------
// loss = mean_b( mean_i( w_b (p - t)^2 ) ) * loss_scale ; dpred = 2 w_b (p - t) / (n B) * loss_scale
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) loss_mse_kernel(const __nv_bfloat16* __restrict__ pred,
                                                               const __nv_bfloat16* __restrict__ target,
                                                               const float* __restrict__ weight, float loss_scale,
                                                               __nv_bfloat16* __restrict__ dpred,
                                                               float* __restrict__ partial, int B,
                                                               long long per_sample) {
    const long long total8 = (long long)B * per_sample / 8;
    float acc = 0.f;
    const float inv = loss_scale / ((float)per_sample * (float)B);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const long long e0 = i * 8;
        const int b = (int)(e0 / per_sample);
        const float w = weight ? weight[b] : 1.f;
        float p[8], t[8], g[8];
        unpack8(ldg16(pred + e0), p);
        unpack8(ldg16(target + e0), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float d = p[e] - t[e];
            acc += w * d * d;
            g[e] = 2.f * w * d * inv;
        }
        if (dpred) stg16(dpred + e0, pack8(g));
    }
    float2 r = block_sum2(acc, 0.f);
    if (threadIdx.x == 0) partial[blockIdx.x] = r.x * inv;
}
__global__ void final_sum_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, int accumulate) {
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += partial[i];
    float2 r = block_sum2(a, 0.f);
    if (threadIdx.x == 0) *out = accumulate ? (*out + r.x) : r.x;
}

__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int n) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 128) return;
    const int i = idx % 128, r = idx / 128;
    const float freq = expf(-9.210340371976184f * (float)i / 128.0f);
    const float a = t[r] * freq;
    out[(long long)r * 256 + i] = __float2bfloat16_rn(cosf(a));
    out[(long long)r * 256 + 128 + i] = __float2bfloat16_rn(sinf(a));
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n,
                                     float scale) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(src + i);
        uint2 o = make_uint2(pack_bf16x2(v.x * scale, v.y * scale), pack_bf16x2(v.z * scale, v.w * scale));
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i] * scale);
    }
}

__global__ void __launch_bounds__(ROW_THREADS) sumsq_kernel(const float* __restrict__ x, long long n,
                                                            float* __restrict__ partial) {
    float acc = 0.f;
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 * 4; i < n; ++i) acc += x[i] * x[i];
    float2 r = block_sum2(acc, 0.f);
    if (threadIdx.x == 0) partial[blockIdx.x] = r.x;
}

// clip (utils/torch.py:99-161: coef = min(1, max_norm / (norm + 1e-6))) + AdamW (torch.optim.AdamW math) + zero grad
__device__ __forceinline__ void adamw_one(float& pi, float& gi_io, float& mi, float& vi, float coef, float lr, float b1,
                                          float b2, float eps, float wd, float bc1, float bc2_sqrt) {
    const float gi = gi_io * coef;
    pi *= (1.f - lr * wd);
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    gi_io = 0.f;
}

// four elements per thread as 16-byte accesses (7 streams of 4 B per element: the kernel is pure HBM traffic, and with one
// element per thread it ran at ~60 % of the copy bandwidth); n4 = n / 4 vectors, the < 4-element tail goes to the last thread
__global__ void __launch_bounds__(256) adamw_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, const float* __restrict__ sumsq,
                                                         float max_norm, float lr, float b1, float b2, float eps, float wd,
                                                         float bc1, float bc2_sqrt, float grad_div) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n4 = n >> 2;
    float coef = grad_div;
    if (max_norm > 0.f) {
        float norm = sqrtf(*sumsq) * grad_div;
        coef *= fminf(1.f, max_norm / (norm + 1e-6f));
    }
    if (i < n4) {
        float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        adamw_one(pp.x, gg.x, mm.x, vv.x, coef, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        adamw_one(pp.y, gg.y, mm.y, vv.y, coef, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        adamw_one(pp.z, gg.z, mm.z, vv.z, coef, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        adamw_one(pp.w, gg.w, mm.w, vv.w, coef, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(g)[i] = gg;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (i == n4) {
        for (long long j = n4 << 2; j < n; ++j) adamw_one(p[j], g[j], m[j], v[j], coef, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
    }
}

}  // namespace b2d

using namespace b2d;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

// instantiate the row kernels for 1..4 chunks of 2048 columns (registers scale with the chunk count)
#define ROW_DISPATCH(D_, KERNEL, GRID, ...)                                             \
    do {                                                                                \
        const int nch__ = ((D_) + 8 * ROW_THREADS - 1) / (8 * ROW_THREADS);             \
        if (nch__ <= 1) launch_k(KERNEL<1>, dim3(GRID), dim3(ROW_THREADS), 0, STREAM, __VA_ARGS__);       \
        else if (nch__ == 2) launch_k(KERNEL<2>, dim3(GRID), dim3(ROW_THREADS), 0, STREAM, __VA_ARGS__);  \
        else launch_k(KERNEL<4>, dim3(GRID), dim3(ROW_THREADS), 0, STREAM, __VA_ARGS__);                  \
    } while (0)

static int check_rowop(int rows, int D, int rps) {
    if (rows <= 0 || D <= 0 || rps <= 0) return set_error(B2D_ERR_SHAPE, "rows/D/rows_per_sample must be positive");
    if (D % 8 != 0 || D > 8 * ROW_THREADS * MAX_CHUNKS) return set_error(B2D_ERR_SHAPE, "D=%d must be a multiple of 8 and <= %d", D, 8 * ROW_THREADS * MAX_CHUNKS);
    return 0;
}

extern "C" int b2d_norm_modulate_fwd(const void* x, void* y, const void* shift_tab, const void* shift_emb,
                                     const void* scale_tab, const void* scale_emb, int64_t emb_stride, int32_t rows,
                                     int32_t D, int32_t rows_per_sample, float eps, int32_t layer_norm, void* stream) {
    B2D_BIND(x);
    if (int rc = check_rowop(rows, D, rows_per_sample)) return rc;
    ROW_DISPATCH(D, norm_modulate_fwd_kernel, rows,
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (const __nv_bfloat16*)shift_tab, (const __nv_bfloat16*)shift_emb,
        (const __nv_bfloat16*)scale_tab, (const __nv_bfloat16*)scale_emb, emb_stride, D, rows_per_sample, eps, layer_norm);
    B2D_CHECK_LAUNCH("norm_modulate_fwd");
    return 0;
}

extern "C" int b2d_norm_modulate_bwd(const void* dy, const void* x, const void* dx_in, void* dx_out,
                                     const void* scale_tab, const void* scale_emb, const void* gate2_tab,
                                     const void* gate2_emb, void* out2, int64_t emb_stride, int32_t rows, int32_t D,
                                     int32_t rows_per_sample, float eps, int32_t layer_norm, void* stream) {
    B2D_BIND(dy);
    if (int rc = check_rowop(rows, D, rows_per_sample)) return rc;
    ROW_DISPATCH(D, norm_modulate_bwd_kernel, rows,
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dx_in, (__nv_bfloat16*)dx_out,
        (const __nv_bfloat16*)scale_tab, (const __nv_bfloat16*)scale_emb, (const __nv_bfloat16*)gate2_tab,
        (const __nv_bfloat16*)gate2_emb, (__nv_bfloat16*)out2, emb_stride, D, rows_per_sample, eps, layer_norm);
    B2D_CHECK_LAUNCH("norm_modulate_bwd");
    return 0;
}

extern "C" int b2d_colscale(const void* x, void* out, const void* tab, const void* emb, int64_t emb_stride,
                            int32_t rows, int32_t D, int32_t rows_per_sample, void* stream) {
    B2D_BIND(x);
    if (D % 8) return set_error(B2D_ERR_SHAPE, "colscale: D %% 8");
    long long total8 = (long long)rows * D / 8;
    launch_k(colscale_kernel, dim3((unsigned)((total8 + 255) / 256)), dim3(256), 0, STREAM, (const __nv_bfloat16*)x,
             (__nv_bfloat16*)out, (const __nv_bfloat16*)tab, (const __nv_bfloat16*)emb, emb_stride, total8, D,
             rows_per_sample);
    B2D_CHECK_LAUNCH("colscale");
    return 0;
}

static int launch_qkv_fwd(const void* src, int64_t ld, int64_t col_off, const QkvSegArgs& a, const void* cos,
                          const void* sin, int B, int S, int H, float eps, void* stream) {
    if (int rc = check_rowop(B * S, H * 64, S)) return rc;
    if ((ld % 8) || (col_off % 8)) return set_error(B2D_ERR_ALIGN, "qkv_norm_rope: ld/col_off must be multiples of 8");
    if (a.nseg < 1 || a.nseg > 3) return set_error(B2D_ERR_SHAPE, "qkv_norm_rope: 1..3 segments");
    if (a.rope_mask && (cos == nullptr || sin == nullptr)) return set_error(B2D_ERR_SHAPE, "qkv_norm_rope: rope needs tables");
    ROW_DISPATCH(H * 64, qkv_norm_rope_fwd_kernel, B * S, (const __nv_bfloat16*)src, ld, col_off, a, (const float*)cos,
                 (const float*)sin, S, H, eps);
    B2D_CHECK_LAUNCH("qkv_norm_rope_fwd");
    return 0;
}

static int launch_qkv_bwd(const void* x, int64_t ld, int64_t col_off, const QkvSegArgs& a, const void* cos,
                          const void* sin, void* dx, int64_t ld_dx, int64_t dx_col_off, int B, int S, int H, float eps,
                          void* stream) {
    if (int rc = check_rowop(B * S, H * 64, S)) return rc;
    if ((ld % 8) || (col_off % 8) || (ld_dx % 8) || (dx_col_off % 8))
        return set_error(B2D_ERR_ALIGN, "qkv_norm_rope_bwd: ld/col_off must be multiples of 8");
    if (a.nseg < 1 || a.nseg > 3) return set_error(B2D_ERR_SHAPE, "qkv_norm_rope_bwd: 1..3 segments");
    if (a.rope_mask && (cos == nullptr || sin == nullptr)) return set_error(B2D_ERR_SHAPE, "qkv_norm_rope_bwd: rope needs tables");
    ROW_DISPATCH(H * 64, qkv_norm_rope_bwd_kernel, B * S, (const __nv_bfloat16*)x, ld, col_off, a, (const float*)cos,
                 (const float*)sin, (__nv_bfloat16*)dx, ld_dx, dx_col_off, S, H, eps);
    B2D_CHECK_LAUNCH("qkv_norm_rope_bwd");
    return 0;
}

extern "C" int b2d_qknorm_rope_fwd(const void* src, int64_t ld, int64_t col_off, const void* weight, const void* cos,
                                   const void* sin, void* dst, int32_t B, int32_t S, int32_t H, int32_t norm,
                                   float eps, void* stream) {
    B2D_BIND(src);
    QkvSegArgs a = {};
    a.nseg = 1;
    a.w[0] = norm ? (const __nv_bfloat16*)weight : nullptr;
    if (norm && weight == nullptr) return set_error(B2D_ERR_SHAPE, "qknorm_rope: norm needs a weight");
    a.dst[0] = (__nv_bfloat16*)dst;
    a.rope_mask = cos != nullptr ? 1 : 0;
    return launch_qkv_fwd(src, ld, col_off, a, cos, sin, B, S, H, eps, stream);
}

extern "C" int b2d_qknorm_rope_bwd(const void* dsrc_heads, const void* x, int64_t ld, int64_t col_off,
                                   const void* weight, const void* cos, const void* sin, void* dx, int64_t ld_dx,
                                   int64_t dx_col_off, int32_t B, int32_t S, int32_t H, int32_t norm, float eps,
                                   void* stream) {
    B2D_BIND(dsrc_heads);
    QkvSegArgs a = {};
    a.nseg = 1;
    a.w[0] = norm ? (const __nv_bfloat16*)weight : nullptr;
    if (norm && weight == nullptr) return set_error(B2D_ERR_SHAPE, "qknorm_rope_bwd: norm needs a weight");
    a.dst[0] = (__nv_bfloat16*)const_cast<void*>(dsrc_heads);
    a.rope_mask = cos != nullptr ? 1 : 0;
    return launch_qkv_bwd(x, ld, col_off, a, cos, sin, dx, ld_dx, dx_col_off, B, S, H, eps, stream);
}

extern "C" int b2d_qkv_norm_rope_fwd(const void* src, int64_t ld, int64_t col_off, int32_t nseg, const void* w0,
                                     const void* w1, const void* w2, int32_t rope_mask, const void* cos, const void* sin,
                                     void* dst0, void* dst1, void* dst2, int32_t B, int32_t S, int32_t H, float eps,
                                     int32_t rows_per_w, int64_t w_stride, void* stream) {
    B2D_BIND(src);
    QkvSegArgs a = {};
    a.nseg = nseg;
    a.w[0] = (const __nv_bfloat16*)w0; a.w[1] = (const __nv_bfloat16*)w1; a.w[2] = (const __nv_bfloat16*)w2;
    a.dst[0] = (__nv_bfloat16*)dst0; a.dst[1] = (__nv_bfloat16*)dst1; a.dst[2] = (__nv_bfloat16*)dst2;
    a.rope_mask = rope_mask;
    a.rows_per_w = rows_per_w; a.w_stride = w_stride;
    if (rows_per_w < 0 || (w_stride % 8) != 0) return set_error(B2D_ERR_ARG, "qkv_norm_rope: bad weight stacking");
    return launch_qkv_fwd(src, ld, col_off, a, cos, sin, B, S, H, eps, stream);
}

extern "C" int b2d_qkv_norm_rope_bwd(const void* dy0, const void* dy1, const void* dy2, const void* x, int64_t ld,
                                     int64_t col_off, int32_t nseg, const void* w0, const void* w1, const void* w2,
                                     int32_t rope_mask, const void* cos, const void* sin, void* dx, int64_t ld_dx,
                                     int64_t dx_col_off, int32_t B, int32_t S, int32_t H, float eps, int32_t rows_per_w,
                                     int64_t w_stride, void* stream) {
    B2D_BIND(dy0);
    QkvSegArgs a = {};
    a.nseg = nseg;
    a.w[0] = (const __nv_bfloat16*)w0; a.w[1] = (const __nv_bfloat16*)w1; a.w[2] = (const __nv_bfloat16*)w2;
    a.dst[0] = (__nv_bfloat16*)const_cast<void*>(dy0);
    a.dst[1] = (__nv_bfloat16*)const_cast<void*>(dy1);
    a.dst[2] = (__nv_bfloat16*)const_cast<void*>(dy2);
    a.rope_mask = rope_mask;
    a.rows_per_w = rows_per_w; a.w_stride = w_stride;
    if (rows_per_w < 0 || (w_stride % 8) != 0) return set_error(B2D_ERR_ARG, "qkv_norm_rope_bwd: bad weight stacking");
    return launch_qkv_bwd(x, ld, col_off, a, cos, sin, dx, ld_dx, dx_col_off, B, S, H, eps, stream);
}

extern "C" int b2d_rope_table(float* cos, float* sin, int32_t F, int32_t H, int32_t W, int32_t D, float sf, float sh,
                              float sw, void* stream) {
    B2D_BIND(cos);
    if (D % 2 || D / 6 < 2) return set_error(B2D_ERR_SHAPE, "rope_table: D must be even and >= 12");
    long long n = (long long)F * H * W * (D / 2);
    rope_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(cos, sin, F, H, W, D, sf, sh, sw);
    B2D_CHECK_LAUNCH("rope_table");
    return 0;
}

extern "C" int b2d_prep_noise_pack(const void* latents, const void* noise, const float* mean, const float* std,
                                   const float* sigma, const float* sigma_ff, void* x_t, void* target, int32_t B,
                                   int32_t C, int32_t F, int32_t HW, void* stream) {
    B2D_BIND(latents);
    long long n = (long long)B * C * F * HW;
    if (n <= 0) return set_error(B2D_ERR_SHAPE, "prep: empty");
    prep_noise_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(
        (const __nv_bfloat16*)latents, (const __nv_bfloat16*)noise, mean, std, sigma, sigma_ff, (__nv_bfloat16*)x_t,
        (__nv_bfloat16*)target, B, C, F, HW);
    B2D_CHECK_LAUNCH("prep_noise_pack");
    return 0;
}

constexpr int REDUCE_BLOCKS = 296;

extern "C" int b2d_loss_mse(const void* pred, const void* target, const float* weight, float loss_scale,
                            float* loss_out, void* dpred, float* partial_ws, int32_t B, int64_t per_sample,
                            void* stream) {
    B2D_BIND(pred);
    if (per_sample % 8) return set_error(B2D_ERR_SHAPE, "loss: per_sample %% 8");
    loss_mse_kernel<<<REDUCE_BLOCKS, ROW_THREADS, 0, STREAM>>>((const __nv_bfloat16*)pred, (const __nv_bfloat16*)target,
                                                               weight, loss_scale, (__nv_bfloat16*)dpred, partial_ws, B,
                                                               per_sample);
    B2D_CHECK_LAUNCH("loss_mse");
    final_sum_kernel<<<1, ROW_THREADS, 0, STREAM>>>(partial_ws, REDUCE_BLOCKS, loss_out, 0);
    B2D_CHECK_LAUNCH("loss_final");
    return 0;
}

extern "C" int b2d_timestep_sinusoid(const float* t, void* out, int32_t n, void* stream) {
    B2D_BIND(t);
    timestep_sinusoid_kernel<<<(n * 128 + 255) / 256, 256, 0, STREAM>>>(t, (__nv_bfloat16*)out, n);
    B2D_CHECK_LAUNCH("timestep_sinusoid");
    return 0;
}

extern "C" int b2d_cast_f32_bf16(const float* src, void* dst, int64_t n, float scale, void* stream) {
    B2D_BIND(src);
    if (n <= 0) return 0;
    long long n4 = (n + 3) / 4;
    cast_f32_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, STREAM>>>(src, (__nv_bfloat16*)dst, n, scale);
    B2D_CHECK_LAUNCH("cast_f32_bf16");
    return 0;
}

extern "C" int b2d_sumsq(const float* x, int64_t n, float* out_sumsq, float* partial_ws, void* stream) {
    B2D_BIND(x);
    sumsq_kernel<<<REDUCE_BLOCKS, ROW_THREADS, 0, STREAM>>>(x, n, partial_ws);
    B2D_CHECK_LAUNCH("sumsq");
    final_sum_kernel<<<1, ROW_THREADS, 0, STREAM>>>(partial_ws, REDUCE_BLOCKS, out_sumsq, 1);
    B2D_CHECK_LAUNCH("sumsq_final");
    return 0;
}

extern "C" int b2d_adamw_clip(float* p, float* g, float* m, float* v, int64_t n, const float* sumsq, float max_norm,
                              float lr, float beta1, float beta2, float eps, float wd, int32_t step, float grad_div,
                              void* stream) {
    B2D_BIND(p);
    if (n <= 0) return 0;
    float bc1 = 1.f - powf(beta1, (float)step);
    float bc2s = sqrtf(1.f - powf(beta2, (float)step));
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
         reinterpret_cast<uintptr_t>(v)) & 15)
        return set_error(B2D_ERR_ALIGN, "adamw_clip: p, g, m, v must be 16-byte aligned");
    adamw_clip_kernel<<<(unsigned)((n / 4 + 1 + 255) / 256), 256, 0, STREAM>>>(p, g, m, v, n, sumsq, max_norm, lr, beta1,
                                                                                beta2, eps, wd, bc1, bc2s, grad_div);
    B2D_CHECK_LAUNCH("adamw_clip");
    return 0;
}
