// TF32 tensor-core GEMMs of the train_latent step.  The reference trains with TF32 matmuls
// (train_latent.py:24-25: torch.backends.cuda.matmul.allow_tf32 = True), so this is the reference's own arithmetic: fp32 storage, operands
// read by tcgen05.mma kind::tf32 (10-bit mantissa), fp32 accumulation in TMEM.
//
//   gemm_tc_nt    C[M,N] (+)= A[M,K] B[N,K]^T   both operands K-contiguous: every forward linear layer (B = W [out][in]) and every data
//                 gradient (B = W^T kept as a transposed copy [in][out]).
//   gemm_tc_tn    C[P,Q] (+)= A[R,P]^T B[R,Q]   both operands MN-contiguous: the weight gradient dW = dY^T X, a reduction over the R rows
//                 (up to millions of edges) that is split over the grid and reduced in a fixed order.
//
// One persistent CTA per SM, 320 threads: warp 0 issues TMA loads into a ring of shared-memory stages (SWIZZLE_128B boxes of 32 fp32
// = 128 bytes), warp 1 issues the MMAs (M = N = 128, K = 8 per instruction) into one of two 128-column TMEM accumulators, warps 2-9
// drain the other accumulator: tcgen05.ld -> swizzled staging tile in shared memory -> TMA store (or TMA reduce-add) of the 128 x 128
// block, which also clips the M / N tails.  Tails of K are zero-filled by the TMA loads.
#include <cstring>

#include "model.h"
#include "tc_common.cuh"
#include "train_ops.h"

namespace cb2 {
namespace train {

using namespace tc;

namespace {

constexpr int TM = 128, TN = 128;
constexpr int EPI_WARPS = 8;                               // two per TMEM lane quarter: each drains two of the four 32-column chunks
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int C_STAGE_BYTES = TM * TN * 4;                 // 64 KB staging of one output block (four 128-row x 128-byte sub-tiles)

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::tf32 instruction descriptor: D fp32, A / B tf32 (format 2), M x N, majors (0 = K-major, 1 = MN-major)
__device__ __forceinline__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tmem_ld32u(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor with an explicit layout type (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B: what MN-major 32-bit operands use)
__device__ __forceinline__ uint64_t umma_desc_lt(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout_type << 61);
}

struct GemmTcArgs {
    int tiles_m, tiles_n;       // output blocks
    int k_steps;                // reduction steps of one block (per split)
    int splits;                 // TN only: the R rows are cut into `splits` ranges; block z writes partial z
    int rows_per_split;         // TN only
    int accumulate;             // C += (TMA reduce-add) instead of C =
    int part_rows;              // TN with splits > 1: row pitch (in output rows) between partials inside the partial buffer map
    int tn_sbo, tn_layout;      // TN: descriptor stride between K groups and layout type (experiment switches)
    int fuse;                   // NT only: 0 = C = A B^T; 1 = C = A B^T + bias; 2 = the same and Y = GELU(C) stored through mapY
    int n_cols;                 // NT: N (guards the bias loads of a partial last column block)
};

// MODE 0 = NT (stage: A box {32 k, 128 m}, B box {32 k, 128 n}, K-major both, 32 k per stage, 4 MMAs)
// MODE 1 = TN (stage: A = four boxes {32 p, 64 r}, B = four boxes {32 q, 64 r}, MN-major both, 64 r per stage, 8 MMAs)
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                             const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapY,
                                                             const float* __restrict__ bias, const GemmTcArgs g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int STAGES = MODE == 0 ? 4 : 2;
    constexpr int OP_BYTES = MODE == 0 ? TM * 32 * 4 : 64 * 128 * 4;           // one operand of one stage: 16 KB (NT) / 32 KB (TN)
    constexpr int STAGE_BYTES = 2 * OP_BYTES;
    unsigned char* sOp = smem;
    unsigned char* sC = smem + STAGES * STAGE_BYTES;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sC + C_STAGE_BYTES);          // full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2 * STAGES + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar_full = [&](int s) { return smem_u32(&sBar[s]); };
    auto bar_empty = [&](int s) { return smem_u32(&sBar[STAGES + s]); };
    auto bar_acc_full = [&](int a) { return smem_u32(&sBar[2 * STAGES + a]); };
    auto bar_acc_empty = [&](int a) { return smem_u32(&sBar[2 * STAGES + 2 + a]); };
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full(a), 1); mbar_init(bar_acc_empty(a), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    const int n_blocks = g.tiles_m * g.tiles_n * g.splits;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        uint32_t it = 0;
        for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
            const int z = blk / (g.tiles_m * g.tiles_n), rem = blk - z * g.tiles_m * g.tiles_n;
            const int mt = rem / g.tiles_n, nt = rem - mt * g.tiles_n;
            for (int k = 0; k < g.k_steps; ++k, ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(bar_empty(s), ((it / STAGES) - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(bar_full(s), (uint32_t)STAGE_BYTES);
                    const uint32_t a = smem_u32(sOp + s * STAGE_BYTES), b = a + OP_BYTES;
                    if (MODE == 0) {
                        tma_load_2d(a, &mapA, k * 32, mt * TM, bar_full(s));
                        tma_load_2d(b, &mapB, k * 32, nt * TN, bar_full(s));
                    } else {
                        const int r0 = z * g.rows_per_split + k * 64;
                        for (int c = 0; c < 4; ++c) {
                            tma_load_2d(a + c * (64 * 128), &mapA, mt * TM + c * 32, r0, bar_full(s));
                            tma_load_2d(b + c * (64 * 128), &mapB, nt * TN + c * 32, r0, bar_full(s));
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t IDESC = MODE == 0 ? umma_idesc_tf32(TM, TN, 0, 0) : umma_idesc_tf32(TM, TN, 1, 1);
        uint32_t it = 0, nb = 0;
        for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++nb) {
            const int ab = nb & 1;
            if (nb >= 2) mbar_wait(bar_acc_empty(ab), ((nb >> 1) - 1) & 1);
            tc_fence_after();
            for (int k = 0; k < g.k_steps; ++k, ++it) {
                const int s = it % STAGES;
                mbar_wait(bar_full(s), (it / STAGES) & 1);
                tc_fence_after();
                const uint32_t a = smem_u32(sOp + s * STAGE_BYTES), b = a + OP_BYTES;
                if (elect_one()) {
                    if (MODE == 0) {
                        const uint64_t a0 = umma_desc(a, 16, 1024), b0 = umma_desc(b, 16, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)              // 8 tf32 = 32 bytes along K per instruction, inside the 128-byte swizzle row
                            umma_tf32(tmem_base + (uint32_t)(ab * 128), a0 + (uint64_t)(kk * 2), b0 + (uint64_t)(kk * 2), IDESC, (k > 0 || kk > 0) ? 1u : 0u);
                    } else {
                        // MN-major: 32 elements (128 bytes) contiguous along M / N, successive 32-element chunks 64 * 128 bytes apart (LBO),
                        // reduction rows 128 bytes apart in groups of 4 (SBO = 512); one K = 8 instruction covers 1024 bytes of rows
                        const uint64_t a0 = umma_desc_lt(a, 64 * 128, (uint32_t)g.tn_sbo, (uint32_t)g.tn_layout), b0 = umma_desc_lt(b, 64 * 128, (uint32_t)g.tn_sbo, (uint32_t)g.tn_layout);
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_tf32(tmem_base + (uint32_t)(ab * 128), a0 + (uint64_t)((kk * 1024) >> 4), b0 + (uint64_t)((kk * 1024) >> 4), IDESC,
                                      (k > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(bar_empty(s));
                    if (k == g.k_steps - 1) umma_commit(bar_acc_full(ab));
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
        const int q = warp & 3, r = q * 32 + lane;                      // (a warp reaches the TMEM lanes 32 (warp % 4) .. + 31)
        const int half = (warp - 2) >> 2;                               // which two column chunks this warp drains
        const bool leader = warp == 2 && lane == 0;
        uint32_t nb = 0;
        for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++nb) {
            const int z = blk / (g.tiles_m * g.tiles_n), rem = blk - z * g.tiles_m * g.tiles_n;
            const int mt = rem / g.tiles_n, nt = rem - mt * g.tiles_n;
            const int ab = nb & 1;
            mbar_wait(bar_acc_full(ab), (nb >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 128);
            // fused forward epilogue (NT): pass 0 stores Z = acc + bias, pass 1 re-reads the accumulator and stores Y = GELU(Z) -- the
            // pre-activation and the activation of a Linear + GELU leave the kernel without a second trip through HBM
            const int passes = (MODE == 0 && g.fuse == 2) ? 2 : 1;
#pragma unroll 1
            for (int pass = 0; pass < passes; ++pass) {
#pragma unroll 1
                for (int c = 2 * half; c < 2 * half + 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32u(trow + (uint32_t)(c * 32), v);
                    if (MODE == 0 && g.fuse != 0) {
                        const int col0 = nt * TN + c * 32;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (col0 + 4 * u + 3 < g.n_cols) b = __ldg(reinterpret_cast<const float4*>(bias + col0 + 4 * u));
                            float z0 = __uint_as_float(v[4 * u]) + b.x, z1 = __uint_as_float(v[4 * u + 1]) + b.y;
                            float z2 = __uint_as_float(v[4 * u + 2]) + b.z, z3 = __uint_as_float(v[4 * u + 3]) + b.w;
                            if (pass == 1) { z0 = gelu_erf(z0); z1 = gelu_erf(z1); z2 = gelu_erf(z2); z3 = gelu_erf(z3); }
                            v[4 * u] = __float_as_uint(z0); v[4 * u + 1] = __float_as_uint(z1);
                            v[4 * u + 2] = __float_as_uint(z2); v[4 * u + 3] = __float_as_uint(z3);
                        }
                    }
                    unsigned char* dst = sC + c * (TM * 128) + r * 128;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        *reinterpret_cast<uint4*>(dst + ((u ^ (r & 7)) << 4)) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                }
                fence_async_smem();
                if (pass == passes - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc_empty(ab));           // this accumulator may be overwritten
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");      // the staging block is complete
                if (leader) {
                    const int row0 = (g.splits > 1 ? z * g.part_rows : 0) + mt * TM;
                    const CUtensorMap* out_map = pass == 0 ? &mapC : &mapY;
                    for (int c = 0; c < 4; ++c) {
                        if (g.accumulate) tma_reduce_add_2d(out_map, nt * TN + c * 32, row0, smem_u32(sC + c * (TM * 128)));
                        else tma_store_2d(out_map, nt * TN + c * 32, row0, smem_u32(sC + c * (TM * 128)));
                    }
                    tma_store_commit();
                    tma_store_wait_read();
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");      // ... and has been read out: it may be refilled
            }
        }
        if (leader) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

// [rows, cols] fp32 row-major tensor with pitch `ld` floats, box {32 columns, box_rows}, SWIZZLE_128B
int encode_f32_map(EncodeTiledFn fn, CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows,
                   CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (f32) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld); return 1; }
    return 0;
}

template <int MODE>
size_t smem_bytes() {
    constexpr int STAGES = MODE == 0 ? 4 : 2;
    constexpr int OP_BYTES = MODE == 0 ? TM * 32 * 4 : 64 * 128 * 4;
    return (size_t)STAGES * 2 * OP_BYTES + C_STAGE_BYTES + (2 * STAGES + 4) * 8 + 16;
}

float* g_part = nullptr;
size_t g_part_bytes = 0;
int g_sms = 0;
EncodeTiledFn g_encode = nullptr;

int prepare() {
    if (g_encode == nullptr) {
        if (get_encode_fn(&g_encode)) return 1;
        int dev = 0;
        CB2_CUDA(cudaGetDevice(&dev));
        CB2_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
        CB2_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<0>()));
        CB2_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<1>()));
    }
    return 0;
}

__global__ void splitk_reduce_tc_kernel(const float* __restrict__ part, int splits, long long MN, int N, long long ldc, float* __restrict__ C, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(long long)z * MN + i];        // fixed order: deterministic
    float* dst = C + (i / N) * ldc + (i % N);
    *dst = accumulate ? *dst + s : s;
}

}  // namespace

bool tc_shape_ok(const float* A, const float* B, const float* C, int M, int N, int K, long long lda, long long ldb, long long ldc, int a_kc, int b_kc) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C);
    if (al % 16 != 0 || lda % 4 != 0 || ldb % 4 != 0 || ldc % 4 != 0) return false;
    if (a_kc && b_kc) return M >= 1024 && N >= 64 && K >= 32;                    // large row counts only: the small GEMMs stay on the SIMT path
    if (!a_kc && !b_kc) return K >= 8192 && M % 128 == 0 && N % 128 == 0;        // weight gradients with full 128 x 128 blocks
    return false;
}

// C[M,N] (+)= A[M,K] B[N,K]^T
int gemm_tc_nt(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb, long long ldc, int accumulate, cudaStream_t s) {
    if (int e = prepare()) return e;
    CUtensorMap mA, mB, mC;
    if (encode_f32_map(g_encode, &mA, A, M, K, lda, TM) || encode_f32_map(g_encode, &mB, B, N, K, ldb, TN) || encode_f32_map(g_encode, &mC, C, M, N, ldc, TM)) return 1;
    GemmTcArgs g{(M + TM - 1) / TM, (N + TN - 1) / TN, (K + 31) / 32, 1, 0, accumulate, 0, 0, 0, 0, N};
    const int blocks = g.tiles_m * g.tiles_n;
    gemm_tc_kernel<0><<<blocks < g_sms ? blocks : g_sms, THREADS, smem_bytes<0>(), s>>>(mA, mB, mC, mC, nullptr, g);
    CB2_LAUNCH_CHECK();
    return 0;
}

// Z[M,N] = A[M,K] B[N,K]^T + bias[N];  Y[M,N] = GELU(Z) when Y != nullptr (same pitch as Z)
int gemm_tc_nt_bias_gelu(const float* A, const float* B, const float* bias, float* Z, float* Y, int M, int N, int K, long long lda, long long ldb,
                         long long ldz, cudaStream_t s) {
    if (int e = prepare()) return e;
    CUtensorMap mA, mB, mZ, mY;
    if (encode_f32_map(g_encode, &mA, A, M, K, lda, TM) || encode_f32_map(g_encode, &mB, B, N, K, ldb, TN) || encode_f32_map(g_encode, &mZ, Z, M, N, ldz, TM)) return 1;
    if (Y != nullptr) { if (encode_f32_map(g_encode, &mY, Y, M, N, ldz, TM)) return 1; }
    else mY = mZ;
    GemmTcArgs g{(M + TM - 1) / TM, (N + TN - 1) / TN, (K + 31) / 32, 1, 0, 0, 0, 0, 0, Y != nullptr ? 2 : 1, N};
    const int blocks = g.tiles_m * g.tiles_n;
    gemm_tc_kernel<0><<<blocks < g_sms ? blocks : g_sms, THREADS, smem_bytes<0>(), s>>>(mA, mB, mZ, mY, bias, g);
    CB2_LAUNCH_CHECK();
    return 0;
}

// C[P,Q] (+)= A[R,P]^T B[R,Q]   (P, Q multiples of 128)
int gemm_tc_tn(const float* A, const float* B, float* C, int P, int Q, int R, long long lda, long long ldb, long long ldc, int accumulate, cudaStream_t s) {
    if (int e = prepare()) return e;
    const int tiles = (P / TM) * (Q / TN);
    int splits = (2 * g_sms) / tiles;
    const int max_splits = (R + 2047) / 2048;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rows_per_split = ((R + splits - 1) / splits + 63) / 64 * 64;
    splits = (R + rows_per_split - 1) / rows_per_split;
    // rows beyond R inside the last split are zero-filled by the loads; rows of the NEXT split must not be read twice: the row box of a
    // split ends exactly at its range because rows_per_split is a multiple of the 64-row stage
    CUtensorMap mA, mB, mC;
    float* out = C;
    long long ld_out = ldc;
    if (splits > 1) {
        const size_t need = (size_t)splits * P * Q * sizeof(float);
        if (need > g_part_bytes) {
            if (g_part) cudaFree(g_part);
            CB2_CUDA(cudaMalloc(&g_part, need));
            g_part_bytes = need;
        }
        out = g_part;
        ld_out = Q;
    }
    // MN-major 32-bit operands use the 32-byte-atom form of the 128-byte swizzle on both sides: TMA SWIZZLE_128B_ATOM_32B writes it, the
    // descriptor names it (layout type 1 = SWIZZLE_128B_BASE32B) with 4-row groups 512 bytes apart.  (Measured: the plain 16-byte-atom
    // SWIZZLE_128B / layout type 2 pairing that serves the 16-bit MN-major operand of the sampling kernels gives wrong sums here.)
    const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    const int tn_layout = 1, tn_sbo = 512;
    if (encode_f32_map(g_encode, &mA, A, R, P, lda, 64, swz) || encode_f32_map(g_encode, &mB, B, R, Q, ldb, 64, swz) ||
        encode_f32_map(g_encode, &mC, out, splits > 1 ? (long long)splits * P : P, Q, ld_out, TM)) return 1;
    GemmTcArgs g{P / TM, Q / TN, rows_per_split / 64, splits, rows_per_split, splits > 1 ? 0 : accumulate, P, tn_sbo, tn_layout, 0, Q};
    const int blocks = tiles * splits;
    gemm_tc_kernel<1><<<blocks < g_sms ? blocks : g_sms, THREADS, smem_bytes<1>(), s>>>(mA, mB, mC, mC, nullptr, g);
    CB2_LAUNCH_CHECK();
    if (splits > 1) {
        const long long MN = (long long)P * Q;
        splitk_reduce_tc_kernel<<<(unsigned)((MN + 255) / 256), 256, 0, s>>>(g_part, splits, MN, Q, ldc, C, accumulate);
        CB2_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace train
}  // namespace cb2
