// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a, next to MUFU.EX2 throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
// One CTA per SM; W warps (multiple of 4) each issue ITERS x { tcgen05.ld.32x32b.x32 ; wait } on their lane quarter.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: ld + wait each; 1: two lds per wait; 2: MUFU only (32 ex2 per iteration); 3: ld overlapped with MUFU
__global__ void __launch_bounds__(512, 1) k(long long* out_clk, float* sink, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32], q[32];
    float acc = 0.f, x = threadIdx.x * 1e-3f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t a = base + ((i * 64) & 448);
        if (MODE == 0 || MODE == 1 || MODE == 3) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(a) : "memory");
        }
        if (MODE == 1) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
                  "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
                  "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
                  "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                : "r"(a + 32) : "memory");
        }
        if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                float y;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x + e * 0.01f));
                acc += y;
            }
            x += 1e-4f;
        }
        if (MODE != 2) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 32; e += 8) acc += __uint_as_float(r[e]);
            if (MODE == 1) acc += __uint_as_float(q[5]);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out_clk[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name, int warps, int iters) {
    long long* d; float* s;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
    k<MODE><<<148, warps * 32>>>(d, s, iters);
    k<MODE><<<148, warps * 32>>>(d, s, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    double clk = (double)h[0];
    const int lds = MODE == 1 ? 2 : (MODE == 2 ? 0 : 1);
    double bytes = (double)warps * iters * lds * 4096.0;
    double exps = (MODE == 2 || MODE == 3) ? (double)warps * iters * 32 * 32 : 0;
    printf("%-34s warps=%2d  clk/iter=%7.1f  tmem B/clk/SM=%7.1f  ex2/clk/SM=%6.2f  %s\n", name, warps, clk / iters, bytes / clk, exps / clk,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main() {
    for (int w : {4, 8, 12, 16}) {
        run<0>("ld.x32 + wait", w, 2000);
        run<1>("2 x ld.x32 + wait", w, 2000);
        run<2>("32 x ex2 only", w, 2000);
        run<3>("ld.x32 overlapped with 32 x ex2", w, 2000);
    }
    return 0;
}
