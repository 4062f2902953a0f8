"""Builds tools/micro/libb2d_trace.so: a copy of the product library with clock64() probes patched into the attention
kernels (development aid only - the product sources carry no trace code).  One thread (CTA 0 of each kernel, first
consumer warp, lane 0) records, per key/query tile: loop top, after the S-ready wait, after tcgen05.wait::ld, before the
P stores, after the barrier arrive.  `b2d_trace_read(buf)` copies the 4096-entry trace out.
   python tools/micro/make_trace_build.py && python tools/micro/attn_trace.py"""
import os
import shutil
import sys
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "finetrainers_b200", "csrc")
DST = os.path.join(HERE, "trace_build")
shutil.rmtree(DST, ignore_errors=True)
os.makedirs(DST)
for f in os.listdir(SRC):
    if f.endswith((".cu", ".cuh", ".h")):
        shutil.copy(os.path.join(SRC, f), DST)
p = os.path.join(DST, "b2d_attn.cu")
s = open(p).read()
s = s.replace('#include "b2d_internal.h"', '#include "../../../finetrainers_b200/csrc/b2d_internal.h"', 1) if False else s
s = s.replace("namespace b2d {\n", "namespace b2d {\n__device__ long long g_trace[16384];\n#define TR(slot) do { if (trace_on) g_trace[(slot)] = clock64(); } while (0)\n", 1)


def must(a, b, count=1):
    global s
    assert a in s, a[:60]
    s = s.replace(a, b, count)


# ---- forward: softmax warps
must("""        for (int j = 0; j < n_kv; ++j) {
            const uint32_t tS = tmem + (j & 1) * 64 + lane_off;
            mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
            tc_fence_after();""",
     """        const bool trace_on = blockIdx.x == 3 && blockIdx.y == 5 && warp == 2 && lane == 0;
        for (int j = 0; j < n_kv; ++j) {
            const uint32_t tS = tmem + (j & 1) * 64 + lane_off;
            TR(j * 8 + 0);
            mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            TR(j * 8 + 1);""")
must("""                tmem_ld32(tS, v0);
                tmem_ld32(tS + 32, v1);
                tmem_ld_wait();
                if (j == 0) {  // first tile""",
     """                tmem_ld32(tS, v0);
                tmem_ld32(tS + 32, v1);
                tmem_ld_wait();
                TR(j * 8 + 2);
                if (j == 0) {  // first tile""")
must("""            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[j & 1]);
        }
        // all P V must have retired""",
     """            TR(j * 8 + 3);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[j & 1]);
            TR(j * 8 + 4);
        }
        // all P V must have retired""")
# ---- backward: consumer warps (dKV and dQ instantiations share the code; slot base 4096 + DKV * 2048)
must("""        for (int it = wg; it < n_y; it += PP_NWG) {
            const int i = y0 + it;""",
     """        const bool trace_on = blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 0 && (warp == 2 || warp == 6) && lane == 0;
        const int tb = 4096 + (DKV ? 2048 : 0) + (warp == 6 ? 1024 : 0);
        for (int it = wg; it < n_y; it += PP_NWG) {
            const int i = y0 + it;
            TR(tb + (it / PP_NWG) * 8 + 0);""")
must("""            mbar_wait(&s_full[kb], (uint32_t)((it / PP_NBUF) & 1));
            tc_fence_after();
            const uint32_t aCA = smem_u32(cA), aCD = smem_u32(cD);""",
     """            TR(tb + (it / PP_NWG) * 8 + 1);
            mbar_wait(&s_full[kb], (uint32_t)((it / PP_NBUF) & 1));
            tc_fence_after();
            TR(tb + (it / PP_NWG) * 8 + 2);
            const uint32_t aCA = smem_u32(cA), aCD = smem_u32(cD);""")
must("""            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&ds_full[wg]);""",
     """            TR(tb + (it / PP_NWG) * 8 + 3);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&ds_full[wg]);
            TR(tb + (it / PP_NWG) * 8 + 4);""")
# ---- backward: the MMA-issuing thread (slots 8192 + DKV * 2048 + it * 4: before / after the ds_full wait, after the
# accumulation GEMMs + commit, after the y_full wait of the S/dP issue that follows) and CTA-level clock64 / globaltimer
# pairs (slots 16000 + DKV * 8): elapsed cycles / elapsed ns = the SM clock the kernel really ran at
must("""            for (int it = 0; it < n_y; ++it) {
                const int st = it % PP_STAGES;
                mbar_wait(&ds_full[it % PP_NWG], (uint32_t)((it / PP_NWG) & 1));  // consumer finished tile it: P^T/dS^T in TMEM
                tc_fence_after();""",
     """            const bool trace_on = blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 0;
            const int mb = 8192 + (DKV ? 2048 : 0);
            for (int it = 0; it < n_y; ++it) {
                const int st = it % PP_STAGES;
                TR(mb + it * 4 + 0);
                mbar_wait(&ds_full[it % PP_NWG], (uint32_t)((it / PP_NWG) & 1));  // consumer finished tile it: P^T/dS^T in TMEM
                tc_fence_after();
                TR(mb + it * 4 + 1);""")
must("""                umma_commit(&y_empty[st]);
                // refill THIS buffer only now: S/dP(it + 4) overwrite the columns the two GEMMs above read (issue order)
                if (it + PP_NBUF < n_y) issue_sdp(it + PP_NBUF);""",
     """                umma_commit(&y_empty[st]);
                TR(mb + it * 4 + 2);
                if (it + PP_NBUF < n_y) {
                    const int st2 = (it + PP_NBUF) % PP_STAGES;
                    mbar_wait(&y_full[st2], (uint32_t)(((it + PP_NBUF) / PP_STAGES) & 1));
                }
                TR(mb + it * 4 + 3);
                if (it + PP_NBUF < n_y) issue_sdp(it + PP_NBUF);""")
must("""    const int n_y = y1 - y0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmX1);""",
     """    const int n_y = y1 - y0;
    if (threadIdx.x == 0 && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 0) {
        long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        g_trace[16000 + (DKV ? 8 : 0)] = clock64();
        g_trace[16001 + (DKV ? 8 : 0)] = gt;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmX1);""")
must("""    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);""",
     """    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 0) {
        long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        g_trace[16002 + (DKV ? 8 : 0)] = clock64();
        g_trace[16003 + (DKV ? 8 : 0)] = gt;
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);""")
s += """
extern "C" int b2d_trace_read(long long* host_dst) {
    return (int)cudaMemcpyFromSymbol(host_dst, b2d::g_trace, sizeof(long long) * 16384);
}
"""
open(p, "w").write(s)
for f in os.listdir(DST):
    if f.endswith((".cu", ".cuh", ".h")):
        t = open(os.path.join(DST, f)).read().replace('"../../include/b2d.h"', '"' + os.path.join(ROOT, "include", "b2d.h") + '"')
        open(os.path.join(DST, f), "w").write(t)
cus = [os.path.join(DST, f) for f in os.listdir(DST) if f.endswith(".cu")]
out = os.path.join(HERE, "libb2d_trace.so")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
       "-diag-suppress", "177", "-shared", "-cudart", "static", "-o", out] + sys.argv[1:] + cus  # extra -D flags
print(" ".join(cmd))
subprocess.check_call(cmd)
print("built", out)
