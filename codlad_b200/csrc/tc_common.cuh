// Inline-PTX building blocks shared by the tcgen05 kernels (edge_tc.cu, node_tc.cu): mbarrier, TMA (bulk tensor
// copies), tcgen05 MMA / TMEM access, UMMA descriptors, and the swizzled operand-tile addressing.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

#include "common.cuh"

namespace cb2 {
namespace tc {

constexpr int TILE_BYTES = 128 * 128 * 2;      // one 128x128 fp16 operand tile (two 64-column SW128 halves)
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr uint32_t SPIN_LIMIT = 1u << 22;      // mbarrier waits trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// debug (CB2_TRAP_DEBUG=1): host-mapped words that receive (barrier address, parity, thread, block) of the wait that timed out
static __device__ unsigned int* g_trap_log = nullptr;
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, bool final) {
    if (g_trap_log != nullptr && blockIdx.x == g_trap_log[0] - 1u + (g_trap_log[0] == 0u ? blockIdx.x + 1u : 0u)) {
        // first block to time out claims the log; every stuck warp of that block records its wait in its own slot
        const unsigned int claimed = atomicCAS(g_trap_log, 0u, blockIdx.x + 1u);
        if (claimed == 0u || claimed == blockIdx.x + 1u) {
            unsigned int* e = g_trap_log + 4 + (threadIdx.x >> 5) * 2;
            e[0] = bar; e[1] = (parity << 16) | (threadIdx.x & 0xffffu);
            __threadfence_system();
        }
    }
    if (final) __trap();
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// Slow path of a wait, kept out of line so that the polling loop, its back-off and the time-out bookkeeping are not replicated
// (and do not hold registers) at every wait of the stage loops.  A waiting warp must not eat the issue slots of the warps doing
// arithmetic on the same scheduler: back off with nanosleep between probes.  The probe count is bounded so that a protocol bug
// traps instead of hanging the GPU.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        __nanosleep(64);
        if (mbar_try(bar, parity)) return;
        if (++spins == SPIN_LIMIT) mbar_timeout(bar, parity, false);      // (debug record; keeps waiting so that the other stuck warps record too)
        if (spins > SPIN_LIMIT + (SPIN_LIMIT >> 1)) mbar_timeout(bar, parity, true);
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major (or MN-major) SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100 version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D fp32, A/B fp16, M x N, majors (0 = K-major, 1 = MN-major)
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of a fully converged warp.  The single-thread instructions (tcgen05.mma / commit, TMA) take their operands from
// uniform registers; issued under elect.sync from warp-uniform control flow they compile to straight-line code, whereas a
// `lane == 0` branch makes the compiler wrap each of them in a per-lane serialisation loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// two x16 loads in flight, one wait (the second load's latency hides behind the first)
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr, float (&v0)[16], float (&v1)[16]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
                   "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr + 16u));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { v0[i] = __uint_as_float(r[i]); v1[i] = __uint_as_float(r[16 + i]); }
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld2_x4(uint32_t taddr, uint32_t col_stride, float (&v)[8]) {     // 4 x (2 columns), one wait
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[2 * i]), "=r"(r[2 * i + 1]) : "r"(taddr + i * col_stride));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void ldg256_coherent(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t u) {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    return __half22float2(h);
}
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// erf-GELU ~= 0.5 x (1 + tanh(x (a + b x^2))), coefficients refitted against the exact form (max |err| 2.7e-4)
__device__ __forceinline__ float gelu_fast(float x) {
    const float x2 = x * x;
    const float u = x * fmaf(x2, 0.03470094f, 0.80015698f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}
// byte offset of the 16-byte chunk `c16` (0..15 across the 256-byte row) of row r inside a swizzled K-major tile
__device__ __forceinline__ uint32_t tile_off(int r, int c16) {
    return (uint32_t)((c16 >> 3) * HALF_BYTES + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}


__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// packed-half helpers of the epilogues
__device__ __forceinline__ uint32_t pack_sat(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// TWICE the erf-GELU of two fp16 values: 2 GELU(x) ~= x (1 + tanh(x (a + b x^2))) (coefficients refitted, see gelu_fast).
// The factor 1/2 is folded into whatever consumes the activation (the fp16 copies of W2 / W12 / W13 / W_out are packed
// pre-multiplied by 0.5, the neighbour-sum indicator holds 0.5), which saves one instruction per element pair.
__device__ __forceinline__ __half2 gelu2_h2(__half2 x) {
#ifdef CB2_X_NOMATH      // timing ablation, never in the product build
    return x;
#endif
    const __half2 x2 = __hmul2(x, x);
    const __half2 pl = __hfma2(x2, __float2half2_rn(0.03470094f), __float2half2_rn(0.80015698f));
    const __half2 u = __hmul2(x, pl);
#ifdef CB2_X_NOTANH      // timing ablation: everything but the two MUFU.TANH + PRMT of the pair
    return __hfma2(x, u, x);
#endif
    uint32_t t;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(as_u32(u)));
    return __hfma2(x, as_h2(t), x);
}


// Programmatic dependent launch: the kernels of the step loop are launched with programmatic stream serialisation, call
// pdl_launch_dependents() at their start (so the NEXT kernel's CTAs can be scheduled on idle SMs and run their prologue --
// barrier init, TMEM allocation, weight TMA) and pdl_wait() before the first access to memory the previous kernel produces.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    static const bool no_pdl = getenv("CB2_NO_PDL") != nullptr;      // debug switch: plain stream order
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = no_pdl ? 0 : 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows, 128] fp16 row-major tensor, box {64 columns, box_rows}, SWIZZLE_128B
inline int encode_rows_map(EncodeTiledFn fn, CUtensorMap* map, const void* base, size_t rows, int box_rows) {
    const cuuint64_t gdim[2] = {128, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {256};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%zu box=%d", (int)r, rows, box_rows); return 1; }
    return 0;
}

inline int get_encode_fn(EncodeTiledFn* out) {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CB2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
    if (!fnp || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return 1; }
    *out = reinterpret_cast<EncodeTiledFn>(fnp);
    return 0;
}

}  // namespace tc

// debug: residency window of one launch (earliest CTA start / latest CTA end, %globaltimer ns) in the plan's "tc_trace" buffer:
// entry 1024 + 2 slot holds max(~start), entry 1025 + 2 slot max(end); zero-initialised
__device__ __forceinline__ void trace_window(unsigned long long* trace, int slot, bool end) {
    if (trace == nullptr) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(trace + 1024 + 2 * slot + (end ? 1 : 0), end ? t : ~t);
}

}  // namespace cb2
