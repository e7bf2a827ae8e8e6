// tcgen05 tier of the per-edge message MLPs (the dominant kernels of the denoiser step).
//
// Same maths as edge_f32.cu (reference models/protein_mpnn_utils.py:240-247, :261-270, :300-307), but the
// chained 128x128 linear layers run on the 5th-generation tensor cores:
//
//   * edge state h_E, activations and weights are fp16 (kind::f16 MMA, fp32 accumulation in TMEM).  fp16 rather
//     than bf16: same tensor rate, 8x finer mantissa, and every operand here is LayerNorm/GELU-bounded.
//   * one persistent CTA per SM over a contiguous, balanced range of tiles: 16 epilogue warps + an MMA warp + a TMA warp (both
//     control warps run warp-wide and issue through elect.sync).  The layer weights stay resident in shared memory
//     (SWIZZLE_128B, K-major, loaded once by TMA); FOUR tiles are in flight (four 32 KB operand buffers, four 128-column TMEM
//     accumulators).  The epilogue warps run the stages round-robin E1(0..3) E2(0..3) E3(0..3), so between two stages of a
//     tile the other three tiles' stages run and cover its MMA / TMA latency; the control warps issue every MMA and TMA the
//     moment its inputs are ready.  A tile = NPT whole nodes (NPT*K <= 128 neighbour rows) of one ensemble member: the
//     neighbour sum is tile-local.
//   * epilogue thread (row r, column quarter cq): the four warps that can address a TMEM lane quarter split the 128
//     accumulator columns of every row, so all 16 warps work on ONE tile stage at a time.
//   * per tile: TMA loads the h_E rows (contiguous in HBM) into a swizzled K-major A tile -> MMA 1 (W?b h_E) ->
//     E1: accumulator from TMEM, + own half Pa[i] (L1 broadcast) + gathered half Pc[j] (fp16, 256-bit loads from L2, issued
//     inside the stage: holding a prefetch in registers costs more than its latency), 2 GELU, fp16 activation back into the
//     same smem tile -> MMA 2 -> E2 (+b, 2 GELU) ->
//        ENC_NODE / DEC : the masked neighbour sum is one more (tiny) MMA: S^T[c, q] = sum_r G[r, c] * Ind[q, r]
//                         with the activation tile reused as an MN-major A operand and a 16-row indicator B (holding the
//                         1/2 of the 2 GELU); slot s is drained (sums read out of TMEM, S stored) by column group s, one
//                         stage later;
//        ENC_EDGE       : MMA 3 (W13) -> the TMA warp re-loads the tile's original rows (residual) -> E3: +residual,
//                         LayerNorm (row statistics exchanged between the four column quarters through drained TMEM
//                         columns), adaLN modulate/gate -> fp16 tile -> TMA store.
//   * epilogue arithmetic is packed fp16 (HFMA2); GELU uses tanh.approx (one MUFU per element) on a refitted 2-term
//     inner polynomial: |err| < 2.8e-4 abs against erf-GELU before the MUFU's own 2^-11 relative error -- the same
//     order as the fp16 rounding of the activation it feeds.  (The fp32 tier uses erff.)
//   * the stage bodies sit at the 96-register cap of an 18-warp CTA: what is kept live across a stage (prefetch buffers, debug
//     timelines, duplicated stage bodies) costs more than it saves -- see DESIGN.md section 4.
#include <cstring>
#include "model.h"
#include "tc_common.cuh"

namespace cb2 {

using namespace tc;

namespace {

constexpr int IND_BYTES = 16 * 128 * 2;        // 16-row indicator operand of the reduction MMA
constexpr int MAX_NPT = 4;

struct TcMaps {
    CUtensorMap in_frame;   // h_E0  [F*L*K, 128] fp16, box {64, K}
    CUtensorMap state;      // h_E   [NB*L*K, 128] fp16, box {64, K}
    CUtensorMap weights;    // packed fp16 weights [n_blocks*128, 128], box {64, 128}
};

struct TcParams {
    int L, K, NPT, tiles_per_member, n_tiles;
    int in_is_frame;                 // layer 0 of the encoder reads h_E0 (indexed by frame)
    int single_frame;                // F == 1: every member uses frame 0 (no frame_of lookup on the metadata path)
    int w_row[3];                    // first row of each weight block in the packed weight tensor
    int n_w;                         // 2 (ENC_NODE / DEC) or 3 (ENC_EDGE)
    const __half* P16;               // [N, 256] fp16: [own half Wa h_V_i + b1 | gathered half Wc h_V_j (+ decoder table)]
    const __half* b2h;               // [128] second-layer bias, fp16
    const __half* b3h;               // [128] third-layer bias (ENC_EDGE), fp16
    const __half* mod16;             // ENC_EDGE: [rows, 3, 256] gate (1 + scale) | gate * shift; member row at b * mod16_stride
    int mod16_stride;
    const __half* res;               // ENC_EDGE: residual source (= the tile's input rows)
    __half* out;                     // ENC_EDGE: h_E rows written by the warpgroup version (the first version stores through TMA)
    const int *lengths, *frame_of, *nbr_idx;
    float* S;                        // [N, 128] neighbour sums (ENC_NODE / DEC)
    unsigned long long* trace;       // debug: stage timestamps of CTA 0 (nullptr = off)
    int trace_slot;                  // debug: launch window slot (trace_window)
    int trace_epi;                   // debug: record the epilogue timeline too (CB2_TRACE_CTRL_ONLY unset)
};

// ---------------------------------------------------------------------------------------------- the kernel
#ifndef CB2_EPI_SUBS
#define CB2_EPI_SUBS 1
#endif
// SUBS 32-column blocks per epilogue thread: 1 = sixteen warps, thread (row, column quarter); 2 = eight warps, thread (row, column half), each
// stage instance walks its two blocks in turn (experiment builds: fewer, fatter warps with 168 registers)
constexpr int SUBS = CB2_EPI_SUBS;
constexpr int NCG = 4 / SUBS;                   // column groups (warps per TMEM lane quarter)
constexpr int EPI_THREADS = 128 * NCG;          // thread (row r, column group cq)
constexpr int CTA_THREADS = EPI_THREADS + 64;   // + two control warps (one lane each): MMA issue, TMA issue
constexpr int NSLOT = 4;                        // tiles in flight per CTA (32 KB operand buffer + 128 TMEM columns each)

// the four warps that hold the column quarters of the same 32 rows (row quarter q): named barrier 1 + q, 128 threads
__device__ __forceinline__ void row_quarter_sync(int q) { asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "n"(32 * NCG) : "memory"); }

// per-thread view of a tile, packed into two registers (four of them are live; they rotate so that the stage code
// exists once):  a = member << 16 | tile-within-member (advanced without divisions) ;  b = member-local neighbour index j
// (raw: nothing waits on its load until the next round's E1), or 0x80000000 for a row outside the tile
struct TileMeta { uint32_t a, b; };
__device__ __forceinline__ int meta_member(const TileMeta& m) { return (int)(m.a >> 16); }
__device__ __forceinline__ int meta_tin(const TileMeta& m) { return (int)(m.a & 0xffffu); }
__device__ __forceinline__ int meta_j(const TileMeta& m) { return (int)(m.b & 0x7fffffffu); }
__device__ __forceinline__ bool meta_row_valid(const TileMeta& m) { return (int)m.b >= 0; }

// MASKED (ENC_NODE / DEC): some rows must be zeroed before the reduction (padded residues, masked neighbours, the unused rows of
// a partial tile).  A template parameter, not a run-time flag: the unmasked build has no trace of the masking in its stage loop.
template <int MODE, bool MASKED>
__global__ void __launch_bounds__(CTA_THREADS, 1) edge_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int N_W = MODE == EDGE_ENC_EDGE ? 3 : 2;
    // layout: [weights N_W x 32 KB][4 tile slots x 32 KB][indicator 4 KB (ENC_NODE / DEC)][barriers]
    unsigned char* sW = smem;
    unsigned char* sT = sW + N_W * TILE_BYTES;
    unsigned char* sAux = sT + 4 * TILE_BYTES;
    unsigned char* sInd = sAux;                                                   // ENC_NODE / DEC
    // barriers: [0] weights; per tile slot g: [1+3g] load (TMA tx), [2+3g] acc (MMA commit), [3+3g] epi (16 warp arrivals); [13+g] go;
    // ENC_EDGE: [17+g] MMA 3 complete (operand tile reusable), [21+g] residual re-load landed
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sAux + (MODE == EDGE_ENC_EDGE ? 0 : IND_BYTES));
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 29);

    const int tid = threadIdx.x;
    if (tid == 0) trace_window(p.trace, p.trace_slot, false);
    const int K = p.K, NPT = p.NPT;
    if ((smem_u32(smem) & 1023u) != 0) __trap();                                   // SWIZZLE_128B atoms repeat every 1 KiB
    pdl_launch_dependents();               // the next kernel of the step may start its prologue on SMs this grid leaves idle

    // ---- one-time setup: barriers, TMEM, indicator operand ----
    if (tid == 0) {
        mbar_init(smem_u32(&sBar[0]), 1);
        for (int g = 0; g < 4; ++g) {
            mbar_init(smem_u32(&sBar[1 + 3 * g]), 1);
            mbar_init(smem_u32(&sBar[2 + 3 * g]), 1);
            mbar_init(smem_u32(&sBar[3 + 3 * g]), EPI_THREADS / 32);          // one arrival per epilogue warp
            mbar_init(smem_u32(&sBar[13 + g]), 1);
            mbar_init(smem_u32(&sBar[17 + g]), 1);
            mbar_init(smem_u32(&sBar[21 + g]), 1);
            mbar_init(smem_u32(&sBar[25 + g]), 4);                     // ENC_NODE / DEC: accumulator drained (the 4 warps of column group g)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid >= EPI_THREADS && tid < EPI_THREADS + 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (MODE != EDGE_ENC_EDGE) {
        // indicator B operand of the reduction MMA: Ind[q][r] = 1/2 if row r belongs to node q (K-major SW128, 16 rows x 128 k)
        for (int t = tid; t < 16 * 16; t += CTA_THREADS) {
            const int q = t >> 4, c16 = t & 15;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r0 = c16 * 8 + e * 2;
                const __half lo = __float2half((q < NPT && r0 / K == q) ? 0.5f : 0.f);        // 0.5: the activations are 2 GELU
                const __half hi = __float2half((q < NPT && (r0 + 1) / K == q) ? 0.5f : 0.f);
                w[e] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
            const uint32_t off = (uint32_t)((c16 >> 3) * (16 * 128) + q * 128 + (((c16 & 7) ^ (q & 7)) << 4));
            *reinterpret_cast<uint4*>(sInd + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    // contiguous, balanced tile range of this CTA (sizes differ by at most one tile), processed NSLOT tiles per round
    const int tile_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
    const int tile_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
    constexpr int tile_stride = 4;
    constexpr uint32_t IDESC_MAIN = umma_idesc(128, 128, 0, 0);
    constexpr uint32_t IDESC_RED = umma_idesc(128, 16, 1, 0);

    if (tid >= EPI_THREADS) {
        // =========================================================================== control warps
        // warp 16 issues every MMA, warp 17 every TMA, in the order the epilogue warps consume them: the epilogue
        // threads never execute issue code and a tile's MMA / TMA latency is covered by the other tiles' epilogues.
        // (MMA and TMA issue are split because a thread's tcgen05.mma stalls behind its own in-flight bulk copies.)
        // Both warps run their control flow with all 32 lanes and issue through elect.sync (see elect_one()).
        auto T_u32 = [&](int g) { return smem_u32(sT + g * TILE_BYTES); };
#ifdef CB2_TRACE_CTRL                                                   // debug builds (tools/dev/build_variant.sh -DCB2_TRACE_CTRL): timeline of CTA 0's
        // control warps (lane 0): MMA warp -> trace[5120..], TMA warp -> trace[5632..]
        unsigned long long* ctr = (p.trace != nullptr && blockIdx.x == 0 && (tid & 31) == 0) ? p.trace + (tid >= EPI_THREADS + 32 ? 5632 : 5120) : nullptr;
        int n_ctr = 0;
        auto cmark = [&](int ev, int g) {
            if (ctr != nullptr && n_ctr < 500) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                ctr[1 + n_ctr] = (t << 8) | (unsigned long long)((ev << 2) | g);
                ctr[0] = (unsigned long long)(++n_ctr);
            }
        };
#else
        auto cmark = [](int, int) {};
#endif
        auto bar_load = [&](int g) { return smem_u32(&sBar[1 + 3 * g]); };
        auto bar_acc = [&](int g) { return smem_u32(&sBar[2 + 3 * g]); };
        auto bar_epi = [&](int g) { return smem_u32(&sBar[3 + 3 * g]); };
        auto bar_go = [&](int g) { return smem_u32(&sBar[13 + g]); };          // slot g's operand tile may be recycled
        auto bar_m3 = [&](int g) { return smem_u32(&sBar[17 + g]); };          // ENC_EDGE: MMA 3 complete
        auto bar_res = [&](int g) { return smem_u32(&sBar[21 + g]); };         // ENC_EDGE: residual rows re-loaded into the tile
        if (tid >= EPI_THREADS + 32) {
            // ------------------------------------------------------------------ TMA warp
            const CUtensorMap* in_map = p.in_is_frame ? &maps.in_frame : &maps.state;
            if (elect_one()) {
                mbar_expect_tx(smem_u32(&sBar[0]), (uint32_t)(N_W * TILE_BYTES));
                for (int m = 0; m < N_W; ++m)
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d(smem_u32(sW + m * TILE_BYTES + h * HALF_BYTES), &maps.weights, h * 64, p.w_row[m], smem_u32(&sBar[0]));
            }
            __syncwarp();
            int nv[4], out_row0[4], in_row0[4];
            auto issue_rows = [&](int g, uint32_t bar) {              // TMA of the tile's h_E rows, one box per node and half
                if (elect_one()) {
                    mbar_expect_tx(bar, (uint32_t)(nv[g] * K * 256));
                    for (int q = 0; q < nv[g]; ++q)
                        for (int h = 0; h < 2; ++h)
                            tma_load_2d(T_u32(g) + h * HALF_BYTES + q * K * 128, in_map, h * 64, in_row0[g] + q * K, bar);
                }
                __syncwarp();
            };
            auto issue_load = [&](int g, int t) {
                const int b = t / p.tiles_per_member;
                const int i0 = (t - b * p.tiles_per_member) * NPT;
                nv[g] = min(NPT, p.L - i0);
                in_row0[g] = ((p.in_is_frame ? __ldg(p.frame_of + b) : b) * p.L + i0) * K;
                out_row0[g] = (b * p.L + i0) * K;
                issue_rows(g, bar_load(g));
            };
            pdl_wait();                                               // h_E is produced by the previous kernels of the step
            for (int g = 0; g < 4; ++g)
                if (tile_begin + g < tile_end) issue_load(g, tile_begin + g);
            uint32_t ph_go = 0;
            for (int t0 = tile_begin; t0 < tile_end; t0 += tile_stride) {
                if (MODE == EDGE_ENC_EDGE) {
                    // residual of the edge update: once MMA 3 has consumed the activation tile, the tile's original h_E rows are
                    // loaded into it again, so E3 reads its residual from shared memory instead of waiting on L2
                    for (int g = 0; g < 4 && t0 + g < tile_end; ++g) { mbar_wait(bar_m3(g), ph_go); issue_rows(g, bar_res(g)); }
                }
                for (int g = 0; g < 4; ++g) {
                    const int t = t0 + g;
                    if (t >= tile_end) break;
                    mbar_wait(bar_go(g), ph_go);                      // reduction MMA complete (ENC_NODE / DEC) or E3 done (ENC_EDGE)
                    cmark(0, g);
                    // (bulk-copy groups belong to the issuing thread: elect.sync picks the same lane every time)
                    if (MODE == EDGE_ENC_EDGE) {
                        if (elect_one()) {
                            for (int q = 0; q < nv[g]; ++q)
                                for (int h = 0; h < 2; ++h)
                                    tma_store_2d(&maps.state, h * 64, out_row0[g] + q * K, T_u32(g) + h * HALF_BYTES + q * K * 128);
                            tma_store_commit();
                        }
                        __syncwarp();
                    }
                    if (t + tile_stride < tile_end) {
                        if (MODE == EDGE_ENC_EDGE) { if (elect_one()) tma_store_wait_read(); __syncwarp(); }   // the store has finished reading the tile
                        issue_load(g, t + tile_stride);
                        cmark(1, g);
                    }
                }
                ph_go ^= 1;
            }
            if (MODE == EDGE_ENC_EDGE) { if (elect_one()) tma_store_wait_all(); __syncwarp(); }
        } else {
            // ------------------------------------------------------------------ MMA warp
            uint32_t ph_epi = 0;
            auto issue_mma = [&](int g, int w_slot) {                 // 128x128x128 GEMM, A = tile g, B = weight slot
                tc_fence_after();
                // descriptors of k-step 0; a k-step advances only the (16-byte granular) start-address field
                const uint64_t a0 = umma_desc(T_u32(g), 16, 1024), b0 = umma_desc(smem_u32(sW + w_slot * TILE_BYTES), 16, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t koff = (uint64_t)(((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4);
                        umma_f16(tmem_base + (uint32_t)(g * 128), a0 + koff, b0 + koff, IDESC_MAIN, k > 0);
                    }
                    umma_commit(bar_acc(g));
                }
                __syncwarp();
            };
            auto issue_reduce = [&](int g) {                          // D[c, q] = sum_r G[r, c] Ind[q, r]
                tc_fence_after();
                const uint64_t a0 = umma_desc(T_u32(g), HALF_BYTES, 1024), b0 = umma_desc(smem_u32(sInd), 16, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)                        // A = G^T: MN-major view, 16 rows (K) per step
                        umma_f16(tmem_base + (uint32_t)(g * 128), a0 + (uint64_t)((k * 16 * 128) >> 4),
                                 b0 + (uint64_t)(((k >> 2) * (16 * 128) + (k & 3) * 32) >> 4), IDESC_RED, k > 0);
                    umma_commit(bar_acc(g));
                    umma_commit(bar_go(g));                            // ... and the operand tile is free once these MMAs complete
                }
                __syncwarp();
            };
            cmark(6, 0);
            mbar_wait(smem_u32(&sBar[0]), 0);                         // weights resident
            cmark(7, 0);
            uint32_t round = 0;
            if (MODE == EDGE_ENC_EDGE) {
                for (int t0 = tile_begin; t0 < tile_end; t0 += tile_stride, ++round) {
                    const int n = min(4, tile_end - t0);                   // live slots of this round
                    const int n_next = max(0, min(4, tile_end - (t0 + tile_stride)));
                    if (round == 0)
                        for (int g = 0; g < n; ++g) { mbar_wait(bar_load(g), 0); issue_mma(g, 0); }            // TMA landed -> MMA 1
                    for (int g = 0; g < n; ++g) { mbar_wait(bar_epi(g), ph_epi); cmark(0, g); issue_mma(g, 1); cmark(1, g); }   // E1 done -> MMA 2
                    ph_epi ^= 1;
                    for (int g = 0; g < n; ++g) {                                                              // E2 done -> MMA 3
                        mbar_wait(bar_epi(g), ph_epi);
                        cmark(2, g);
                        issue_mma(g, 2);
                        if (elect_one()) umma_commit(bar_m3(g));
                        __syncwarp();
                    }
                    ph_epi ^= 1;
                    // E3 done: the tile is complete in shared memory and the accumulator is drained.  The TMA warp stores the tile
                    // and only then reloads the slot, so the slot's next MMA 1 trails by one slot.
                    const uint32_t next_parity = (round + 1) & 1;
                    for (int g = 0; g < n; ++g) {
                        mbar_wait(bar_epi(g), ph_epi);
                        if (elect_one()) mbar_arrive(bar_go(g));                                               // store + recycle
                        __syncwarp();
                        cmark(3, g);
                        if (g >= 1 && g - 1 < n_next) { mbar_wait(bar_load(g - 1), next_parity); cmark(4, g - 1); issue_mma(g - 1, 0); cmark(5, g - 1); }
                    }
                    if (n - 1 < n_next) { mbar_wait(bar_load(n - 1), next_parity); issue_mma(n - 1, 0); }
                    ph_epi ^= 1;
                }
            } else {
                // ENC_NODE / DEC.  A slot's next tile is started as soon as it can be: right after the NEXT slot's reduction has been
                // issued (by then the drain warps have read the slot's sums and its reload, requested when its own reduction
                // completed, has landed) -- not after the round's last reduction, which left the epilogue warps idle at every round
                // boundary.  Slot 3's turn comes after the first MMA 2 of the following round (its drain rides behind E1 of slot 0).
                bool pend3 = false;                                    // slot 3 of the previous round still has to be restarted
                uint32_t pend_round = 0;
                auto restart = [&](int h, uint32_t rnd) {              // next tile of slot h; rnd = round of the tile that just finished
                    mbar_wait(smem_u32(&sBar[25 + h]), rnd & 1);       // reduced sums read out of TMEM
                    cmark(3, h);
                    mbar_wait(bar_load(h), (rnd + 1) & 1);
                    cmark(4, h);
                    issue_mma(h, 0);
                    cmark(5, h);
                };
                for (int t0 = tile_begin; t0 < tile_end; t0 += tile_stride, ++round) {
                    const int n = min(4, tile_end - t0);
                    const int n_next = max(0, min(4, tile_end - (t0 + tile_stride)));
                    if (round == 0)
                        for (int g = 0; g < n; ++g) { mbar_wait(bar_load(g), 0); issue_mma(g, 0); }
                    for (int g = 0; g < n; ++g) {                                                              // E1 done -> MMA 2
                        mbar_wait(bar_epi(g), ph_epi); cmark(0, g); issue_mma(g, 1); cmark(1, g);
                        if (g == 0 && pend3) { restart(3, pend_round); pend3 = false; }
                    }
                    ph_epi ^= 1;
                    for (int g = 0; g < n; ++g) {                                                              // E2 done -> reduction MMA
                        mbar_wait(bar_epi(g), ph_epi);
                        cmark(2, g);
                        issue_reduce(g);
                        if (g >= 1 && g - 1 < n_next) restart(g - 1, round);
                    }
                    ph_epi ^= 1;
                    if (n - 1 < n_next) {
                        if (n == 4) { pend3 = true; pend_round = round; }
                        else restart(n - 1, round);
                    }
                }
            }
        }
    } else {
        // =========================================================================== epilogue threads
        const int warp = tid >> 5, quarter = warp & 3, cq = warp >> 2, r = quarter * 32 + (tid & 31);
        const int c0 = cq * 32 * SUBS;                                 // this thread's first accumulator column (32 SUBS columns in all)
        const int q_of_r = r / K;
        const int len0 = __ldg(p.lengths);                             // length of frame 0 (the only frame of an ensemble plan)
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        // stage hand-off to the control lanes: every thread orders its own smem writes / TMEM reads, the warp converges,
        // one lane arrives (512 arrivals on one mbarrier word would serialise)
        auto stage_done = [&](int s, bool wrote_smem) {
            if (wrote_smem) fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(smem_u32(&sBar[3 + 3 * s]));
        };

        // metadata of a tile for this thread (one global load, consumed a full round later).  (member, tile-in-member) advance
        // by a fixed step per round, so there is no integer division on the per-tile path.
        const int step_b = tile_stride / p.tiles_per_member, step_t = tile_stride - step_b * p.tiles_per_member;
        auto make_meta = [&](int tile, int b, int tin) {
            TileMeta m{((uint32_t)b << 16) | (uint32_t)tin, 0x80000000u};
            if (tile >= tile_end) { m.a = 0u; return m; }              // past the end: keep every derived address in bounds
            const int i0 = tin * NPT;
            const int f = p.single_frame ? 0 : __ldg(p.frame_of + b);    // (dependent load only for multi-frame plans)
            if (q_of_r < min(NPT, p.L - i0))
                m.b = (uint32_t)__ldg(p.nbr_idx + ((size_t)f * p.L + i0 + q_of_r) * K + (r - q_of_r * K));
            return m;
        };
        auto first_meta = [&](int tile) {
            const int b = tile / p.tiles_per_member;
            return make_meta(tile, b, tile - b * p.tiles_per_member);
        };
        auto next_meta = [&](const TileMeta& m, int tile) {               // the slot's tile of the next round
            int b = meta_member(m) + step_b, tin = meta_tin(m) + step_t;
            if (tin >= p.tiles_per_member) { tin -= p.tiles_per_member; ++b; }
            return make_meta(tile, b, tin);
        };
        auto node0_of = [&](const TileMeta& m) { return meta_member(m) * p.L + meta_tin(m) * NPT; };
        auto nv_of = [&](const TileMeta& m) { return min(NPT, p.L - meta_tin(m) * NPT); };
        // neighbour-sum mask of this thread's row (reference: mask_attend = mask_i mask_j; the decoder passes no mask)
        auto row_keep = [&](const TileMeta& m) -> uint32_t {
            if (!meta_row_valid(m)) return 0u;
            if (MODE == EDGE_DEC) return 0xffffffffu;
            const int len = p.single_frame ? len0 : __ldg(p.lengths + __ldg(p.frame_of + meta_member(m)));
            return (meta_tin(m) * NPT + q_of_r < len && meta_j(m) < len) ? 0xffffffffu : 0u;
        };
        auto ld_pc = [&](const TileMeta& m, int col, uint32_t (&pc)[16]) {       // 32 gathered halves of Pc[j] starting at column col
#ifdef CB2_X_NOGATHER
            const __half* src = p.P16 + 128 + col;
#else
            const __half* src = p.P16 + ((size_t)meta_member(m) * p.L + meta_j(m)) * 256 + 128 + col;
#endif
            ldg256(src, *reinterpret_cast<uint32_t(*)[8]>(&pc[0]));
            ldg256(src + 16, *reinterpret_cast<uint32_t(*)[8]>(&pc[8]));
        };
        // packed-half store of 16 consecutive columns [c0 + g16*16, +16) of row r
        auto st16 = [&](unsigned char* T, int col, int g16, const uint32_t (&o)[8]) {
            *reinterpret_cast<uint4*>(T + tile_off(r, (col >> 3) + g16 * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(T + tile_off(r, (col >> 3) + g16 * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
        };

#ifdef CB2_TRACE_EPI                                                    // epilogue timeline of CTA 0 / thread 0 (debug builds only: the marks
        unsigned long long* trace = (p.trace != nullptr && p.trace_epi && blockIdx.x == 0 && tid == 0) ? p.trace : nullptr;      // cost registers and code
        int n_trace = 0;                                                //  in every stage of the hot loop)
        auto mark = [&](int ev, int s) {
            if (trace != nullptr && n_trace < 500) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                trace[1 + n_trace] = (t << 8) | (unsigned long long)((ev << 2) | s);
                trace[0] = (unsigned long long)(++n_trace);
            }
        };
#else
        auto mark = [](int, int) {};
#endif
        TileMeta m0 = first_meta(tile_begin + 0), m1 = first_meta(tile_begin + 1), m2 = first_meta(tile_begin + 2), m3 = first_meta(tile_begin + 3);
        auto rotate = [&]() { const TileMeta t = m0; m0 = m1; m1 = m2; m2 = m3; m3 = t; };
        uint32_t ph = 0;                                                // parity of the acc barriers (all slots advance in lock step)
        uint32_t ph_res = 0;                                            // parity of the residual re-load barriers (one phase per round)
        pdl_wait();                                                     // P16 / S / h_E belong to the previous kernels of the step

        // ENC_NODE / DEC: "E3" (read the reduced sums out of TMEM, store S, release the accumulator, fetch the slot's next
        // metadata) is not a stage of its own: it rides at the tail of the FOLLOWING stage, when the (short) reduction MMA
        // has long completed, so nothing waits on it.  Operates on m3 = the slot processed one stage ago.
        auto drain = [&](int s, int next_tile) {
            const TileMeta m = m3;
            m3 = next_meta(m, next_tile);
            if (cq == s % NCG) {                                         // slot s is drained by column group s (one warp per scheduler): the four
                                                                         // groups share the drains, the other twelve warps run ahead
                mark(8, s);
                mbar_wait(smem_u32(&sBar[2 + 3 * s]), ph ^ 1);           // third commit of the tile (reduction MMA)
                tc_fence_after();
                mark(9, s);
                float s4[4];
                tmem_ld4(tmem_lane - (uint32_t)c0 + (uint32_t)(s * 128), s4);     // lane = output column, 4 columns = nodes of the tile
                const int nv = nv_of(m), node0 = node0_of(m);
#pragma unroll
                for (int q = 0; q < MAX_NPT; ++q)
                    if (q < nv) p.S[((size_t)node0 + q) * 128 + r] = s4[q];
                tc_fence_before();
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(smem_u32(&sBar[25 + s]));   // accumulator drained: the slot's next MMA 1 may start
                mark(10, s);
            }
        };
        bool pending3 = false;                                           // slot 3 of the previous round still has to be drained

        for (int t0 = tile_begin; t0 < tile_end; t0 += tile_stride) {
            const int n = min(NSLOT, tile_end - t0);                    // live slots of this round
            // ================= E1: GELU(acc + Pa[i] + Pc[j]) -> fp16 activation tile
            auto epi1 = [&](int s) {
                unsigned char* T = sT + s * TILE_BYTES;
                uint32_t pc[SUBS][16];                                  // gathered half of this tile's rows (loaded here, not a stage ahead: measured slower,
                uint32_t pa[SUBS][16];                                  //  with 96 and with 168 registers -- DESIGN.md section 4)
                const __half* pa_src = p.P16 + (size_t)(node0_of(m0) + (meta_row_valid(m0) ? q_of_r : 0)) * 256 + c0;   // own half: two nodes per tile, L1-resident
#pragma unroll
                for (int sub = 0; sub < SUBS; ++sub) {
                    ld_pc(m0, c0 + sub * 32, pc[sub]);
                    ldg256(pa_src + sub * 32, *reinterpret_cast<uint32_t(*)[8]>(&pa[sub][0]));
                    ldg256(pa_src + sub * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&pa[sub][8]));
                }
                mark(0, s);
                // The warps that do not drain slot s (cq != s) never wait for the slot's third commit (the reduction MMA), and the phases
                // on either side of it have the same parity: a warp that ran a whole E2 phase ahead of a straggler would sail
                // through the accumulator wait below on the stale "MMA 2 complete" state (seen as a rare dead-lock on large
                // working sets, where TLB misses skew the warps).  The slot's "drained" barrier advances once per round and
                // only after the reduction has completed, so it orders them exactly.
                if (MODE != EDGE_ENC_EDGE && cq != s % NCG && t0 != tile_begin) mbar_wait(smem_u32(&sBar[25 + s]), ph ^ 1);
                mbar_wait(smem_u32(&sBar[2 + 3 * s]), ph);
                tc_fence_after();
                mark(1, s);
#pragma unroll
                for (int sub = 0; sub < SUBS; ++sub) {
                    float acc0[16], acc1[16];
#ifdef CB2_X_NOLDTM                                                     // timing ablations (tools/dev/build_variant.sh), never in the product build
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        asm volatile("mov.b32 %0, %1;" : "=f"(acc0[e]) : "r"(tid + e));
                        asm volatile("mov.b32 %0, %1;" : "=f"(acc1[e]) : "r"(tid - e));
                    }
#else
                    tmem_ld16x2(tmem_lane + (uint32_t)(s * 128 + sub * 32), acc0, acc1);
#endif
#pragma unroll
                    for (int g16 = 0; g16 < 2; ++g16) {
                        const float (&acc)[16] = g16 ? acc1 : acc0;
                        uint32_t o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const __half2 x = __hadd2(__hadd2(as_h2(pack_sat(acc[e * 2], acc[e * 2 + 1])), as_h2(pa[sub][g16 * 8 + e])), as_h2(pc[sub][g16 * 8 + e]));
                            o[e] = as_u32(gelu2_h2(x));
                        }
#ifdef CB2_X_NOSTS
                        if (o[0] == 0x12345678u && o[5] == 0x9abcdef0u) st16(T, c0 + sub * 32, g16, o);
#else
                        st16(T, c0 + sub * 32, g16, o);
#endif
                    }
                }
                mark(2, s);
                stage_done(s, true);
                mark(3, s);
            };
#pragma unroll 1
            for (int s = 0; s < NSLOT; ++s) {
                if (s < n) epi1(s);
                if (MODE != EDGE_ENC_EDGE && s == 0 && pending3) { drain(3, t0 + 3); pending3 = false; }
                rotate();
            }
            ph ^= 1;
            // ================= E2: GELU(acc + b2) (masked rows -> 0 for the reduction) -> fp16 tile
#pragma unroll 1
            for (int s = 0; s < NSLOT; ++s) {
                if (s < n) {
                    unsigned char* T = sT + s * TILE_BYTES;
                    // rows outside the neighbour sum (padded residues, masked neighbours, the unused rows of a partial tile) are zeroed;
                    // when the geometry has none of them (full-length frames, NPT K = 128) the whole step is skipped
                    constexpr bool masked = MODE != EDGE_ENC_EDGE && MASKED;
                    const uint32_t keep = masked ? row_keep(m0) : 0xffffffffu;
                    uint32_t bb[SUBS][16];
#pragma unroll
                    for (int sub = 0; sub < SUBS; ++sub) {
                        ldg256(p.b2h + c0 + sub * 32, *reinterpret_cast<uint32_t(*)[8]>(&bb[sub][0]));
                        ldg256(p.b2h + c0 + sub * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&bb[sub][8]));
                    }
                    mark(4, s);
                    mbar_wait(smem_u32(&sBar[2 + 3 * s]), ph);
                    tc_fence_after();
                    mark(5, s);
#pragma unroll
                    for (int sub = 0; sub < SUBS; ++sub) {
                        float acc0[16], acc1[16];
#ifdef CB2_X_NOLDTM
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            asm volatile("mov.b32 %0, %1;" : "=f"(acc0[e]) : "r"(tid + e));
                            asm volatile("mov.b32 %0, %1;" : "=f"(acc1[e]) : "r"(tid - e));
                        }
#else
                        tmem_ld16x2(tmem_lane + (uint32_t)(s * 128 + sub * 32), acc0, acc1);
#endif
#pragma unroll
                        for (int g16 = 0; g16 < 2; ++g16) {
                            const float (&acc)[16] = g16 ? acc1 : acc0;
                            uint32_t o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const __half2 x = __hadd2(as_h2(pack_sat(acc[e * 2], acc[e * 2 + 1])), as_h2(bb[sub][g16 * 8 + e]));
                                o[e] = as_u32(gelu2_h2(x));
                            }
                            if (masked) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) o[e] &= keep;
                            }
#ifdef CB2_X_NOSTS
                            if (o[0] == 0x12345678u && o[5] == 0x9abcdef0u) st16(T, c0 + sub * 32, g16, o);
#else
                            st16(T, c0 + sub * 32, g16, o);
#endif
                        }
                    }
                    mark(6, s);
                    stage_done(s, true);
                    mark(7, s);
                }
                if (MODE != EDGE_ENC_EDGE && s >= 1 && s - 1 < n) drain(s - 1, t0 + tile_stride + s - 1);
                rotate();
            }
            if (MODE != EDGE_ENC_EDGE) pending3 = n == NSLOT;
            ph ^= 1;
            // ================= E3 (ENC_EDGE): residual + LayerNorm + adaLN; fetch the metadata of the slot's next tile
#pragma unroll 1
            for (int s = 0; s < NSLOT; ++s) {
                if (MODE == EDGE_ENC_EDGE && s < n) {
                    const TileMeta m = m0;
                    mark(8, s);
                    m0 = next_meta(m, t0 + tile_stride + s);
                    {
                        unsigned char* T = sT + s * TILE_BYTES;
                        const int bmem = meta_member(m);
                        mbar_wait(smem_u32(&sBar[2 + 3 * s]), ph);
                        tc_fence_after();
                        mbar_wait(smem_u32(&sBar[21 + s]), ph_res);              // the tile holds the original h_E rows again (residual)
                        // pass A: v = residual + (acc + b13) in packed half, kept in registers; row statistics of those values in fp32
                        float sum = 0.f, sq = 0.f;
                        uint32_t v16[SUBS][16];
#pragma unroll
                        for (int sub = 0; sub < SUBS; ++sub) {
                            const int col = c0 + sub * 32;
                            uint32_t rs[16];
#pragma unroll
                            for (int c16 = 0; c16 < 4; ++c16) {
                                const uint4 t4 = *reinterpret_cast<const uint4*>(T + tile_off(r, (col >> 3) + c16));
                                rs[c16 * 4] = t4.x; rs[c16 * 4 + 1] = t4.y; rs[c16 * 4 + 2] = t4.z; rs[c16 * 4 + 3] = t4.w;
                            }
                            uint32_t b3r[16];
                            ldg256(p.b3h + col, *reinterpret_cast<uint32_t(*)[8]>(&b3r[0]));
                            ldg256(p.b3h + col + 16, *reinterpret_cast<uint32_t(*)[8]>(&b3r[8]));
#pragma unroll
                            for (int g16 = 0; g16 < 2; ++g16) {
                                float acc[16];
                                tmem_ld16(tmem_lane + (uint32_t)(s * 128 + sub * 32 + g16 * 16), acc);
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const __half2 v = __hadd2(as_h2(rs[g16 * 8 + e]), __hadd2(as_h2(pack_sat(acc[e * 2], acc[e * 2 + 1])), as_h2(b3r[g16 * 8 + e])));
                                    const float2 vf = __half22float2(v);
                                    sum += vf.x + vf.y;
                                    sq = fmaf(vf.x, vf.x, fmaf(vf.y, vf.y, sq));
                                    v16[sub][g16 * 8 + e] = as_u32(v);
                                }
                            }
                        }
                        // row statistics: the column groups of a row exchange their partial sums through accumulator
                        // columns this thread has already drained (its own first two) -- tcgen05.st, one barrier, tcgen05.ld;
                        // summed in a fixed order, so the result is deterministic
#ifdef CB2_X_NOXCHG                                                     // timing ablation: no exchange of the row statistics
                        const float tsum = 4.0f * sum, tsq = 4.0f * sq;
#else
                        tmem_st2(tmem_lane + (uint32_t)(s * 128), sum, sq);
                        tc_fence_before();
                        row_quarter_sync(quarter);
                        tc_fence_after();
                        float part[8];
                        if (SUBS == 1) {
                            tmem_ld2_x4(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 128), 32u, part);
                        } else {
                            uint32_t r4[4];
                            const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 128);
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r4[0]), "=r"(r4[1]) : "r"(ta));
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r4[2]), "=r"(r4[3]) : "r"(ta + 64u));
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int i = 0; i < 4; ++i) part[i] = __uint_as_float(r4[i]);
                            part[4] = part[5] = part[6] = part[7] = 0.f;
                        }
                        const float tsum = (part[0] + part[2]) + (part[4] + part[6]), tsq = (part[1] + part[3]) + (part[5] + part[7]);
#endif
                        const float mean = tsum * (1.0f / 128.0f);
                        const float rstd = rsqrtf(fmaxf(tsq * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-6f);
                        const __half2 rstd2 = __float2half2_rn(rstd), nmr2 = __float2half2_rn(-mean * rstd);
                        // pass B (packed half): out = (v rstd - mean rstd) A[c] + B[c],  A = gate (1 + scale), B = gate * shift
#pragma unroll
                        for (int sub = 0; sub < SUBS; ++sub) {
                            const int col = c0 + sub * 32;
                            const __half* mod_row = p.mod16 + (size_t)bmem * p.mod16_stride + col;
#pragma unroll
                            for (int c16 = 0; c16 < 4; ++c16) {
                                const uint4 av = __ldg(reinterpret_cast<const uint4*>(mod_row + c16 * 8));
                                const uint4 bv = __ldg(reinterpret_cast<const uint4*>(mod_row + 128 + c16 * 8));
                                const uint32_t a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
                                uint32_t o[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    o[e] = as_u32(__hfma2(__hfma2(as_h2(v16[sub][c16 * 4 + e]), rstd2, nmr2), as_h2(a4[e]), as_h2(b4[e])));
                                *reinterpret_cast<uint4*>(T + tile_off(r, (col >> 3) + c16)) = make_uint4(o[0], o[1], o[2], o[3]);
                            }
                        }
                        stage_done(s, true);                             // tile complete: the TMA lane stores it and reloads the slot
                        mark(9, s);
                    }
                }
                if (MODE == EDGE_ENC_EDGE) rotate();
            }
            ph ^= 1;
            ph_res ^= 1;
        }
        if (MODE != EDGE_ENC_EDGE && pending3) drain(3, tile_end);        // (ph ^ 1 inside drain = the last round's third parity)
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) trace_window(p.trace, p.trace_slot, true);
    if (tid >= EPI_THREADS && tid < EPI_THREADS + 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace
}  // namespace cb2
// The warpgroup-per-tile variant of these kernels (edge_wg.cuh) was built and measured in round 2 and is slower at every shape
// ("warpgroup variant"); it is compiled only into experiment builds (tools/dev/build_variant.sh -DCB2_WITH_WG,
// selected at run time with CB2_EDGE_WG=1), never into the product library.
namespace cb2 {
namespace {
#ifdef CB2_WITH_WG
#include "edge_wg.cuh"
#endif

size_t tc_smem_bytes(int mode) {
    const int n_w = mode == EDGE_ENC_EDGE ? 3 : 2;
    return (size_t)(n_w + 4) * TILE_BYTES + (mode == EDGE_ENC_EDGE ? 0 : IND_BYTES) + 29 * 8 + 16;
}

}  // namespace

unsigned int* g_trap_host = nullptr;       // CB2_TRAP_DEBUG: host view of the time-out record of this file's kernels

const unsigned int* edge_tc_trap_log() { return g_trap_host; }

int edge_tc_prepare(Plan& p) {
    if (p.K > 128 || p.K < 1) { set_error("edge_tc: K=%d unsupported", p.K); return 1; }
    EncodeTiledFn fn = nullptr;
    if (get_encode_fn(&fn)) return 1;
    TcMaps* m = new TcMaps();
    p.tmaps = m;
    if (encode_rows_map(fn, &m->in_frame, p.hE0, (size_t)p.F * p.L * p.K, p.K)) return 1;
    if (encode_rows_map(fn, &m->state, p.hE, (size_t)p.NB * p.L * p.K, p.K)) return 1;
    if (encode_rows_map(fn, &m->weights, p.model->dev_f16, (size_t)p.model->n_f16_blocks * 128, 128)) return 1;
    if (getenv("CB2_TRAP_DEBUG") != nullptr && g_trap_host == nullptr) {
        unsigned int* dptr = nullptr;
        CB2_CUDA(cudaHostAlloc(&g_trap_host, 256, cudaHostAllocMapped));
        memset(g_trap_host, 0, 256);
        CB2_CUDA(cudaHostGetDevicePointer(&dptr, g_trap_host, 0));
        CB2_CUDA(cudaMemcpyToSymbol(tc::g_trap_log, &dptr, sizeof(dptr)));
    }
    int dev = 0;
    CB2_CUDA(cudaGetDevice(&dev));
    CB2_CUDA(cudaDeviceGetAttribute(&p.num_sms, cudaDevAttrMultiProcessorCount, dev));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<EDGE_ENC_NODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(EDGE_ENC_NODE)));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<EDGE_ENC_NODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(EDGE_ENC_NODE)));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<EDGE_DEC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(EDGE_DEC)));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<EDGE_DEC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(EDGE_DEC)));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<EDGE_ENC_EDGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(EDGE_ENC_EDGE)));
#ifdef CB2_WITH_WG
    CB2_CUDA(cudaFuncSetAttribute(wg::edge_wg_kernel<EDGE_ENC_NODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::wg_smem_bytes(EDGE_ENC_NODE)));
    CB2_CUDA(cudaFuncSetAttribute(wg::edge_wg_kernel<EDGE_ENC_NODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::wg_smem_bytes(EDGE_ENC_NODE)));
    CB2_CUDA(cudaFuncSetAttribute(wg::edge_wg_kernel<EDGE_DEC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::wg_smem_bytes(EDGE_DEC)));
    CB2_CUDA(cudaFuncSetAttribute(wg::edge_wg_kernel<EDGE_DEC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::wg_smem_bytes(EDGE_DEC)));
    CB2_CUDA(cudaFuncSetAttribute(wg::edge_wg_kernel<EDGE_ENC_EDGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::wg_smem_bytes(EDGE_ENC_EDGE)));
#endif
    return 0;
}

void edge_tc_release(Plan& p) {
    delete reinterpret_cast<TcMaps*>(p.tmaps);
    p.tmaps = nullptr;
}

int launch_edge_tc(Plan& p, int mode, int layer, const float* mod_base, int mod_stride_b, cudaStream_t s) {
    const DenoiserModel& m = *p.model;
    const TcMaps& maps = *reinterpret_cast<const TcMaps*>(p.tmaps);
    TcParams tp{};
    tp.L = p.L; tp.K = p.K;
    if (p.NB > 65535 || p.L > 65535) { set_error("edge_tc: NB=%d / L=%d exceed the 16-bit tile metadata fields", p.NB, p.L); return 1; }
    tp.NPT = 128 / p.K < MAX_NPT ? 128 / p.K : MAX_NPT;
    if (tp.NPT < 1 || p.K % 8 != 0) tp.NPT = 1;      // node row blocks must start on a swizzle-atom (8-row) boundary
    tp.tiles_per_member = (p.L + tp.NPT - 1) / tp.NPT;
    tp.n_tiles = tp.tiles_per_member * p.NB;
    tp.lengths = p.lengths; tp.frame_of = p.frame_of; tp.nbr_idx = p.nbr_idx; tp.S = p.S;
    tp.trace = p.tc_trace; tp.trace_slot = p.launches & 2047;
    { static const bool ctrl_only = getenv("CB2_TRACE_CTRL_ONLY") != nullptr; tp.trace_epi = ctrl_only ? 0 : 1; }
    const bool first = (layer == 0 && mode != EDGE_DEC);
    tp.in_is_frame = first ? 1 : 0;
    tp.single_frame = p.F == 1 ? 1 : 0;
    const bool masked = tp.NPT * p.K < 128 || p.L % tp.NPT != 0 || (mode != EDGE_DEC && !p.all_full);
    auto row_of = [&](const __half* w) { return (int)((w - m.dev_f16) / 128); };
    if (mode == EDGE_ENC_NODE) {
        const EncLayerW& e = m.enc[layer];
        tp.P16 = p.P16[0]; tp.n_w = 2;
        tp.w_row[0] = row_of(e.W1b_h); tp.w_row[1] = row_of(e.W2_h); tp.b2h = e.b2_16;
    } else if (mode == EDGE_ENC_EDGE) {
        const EncLayerW& e = m.enc[layer];
        tp.P16 = p.P16[1]; tp.n_w = 3;
        tp.w_row[0] = row_of(e.W11b_h); tp.w_row[1] = row_of(e.W12_h); tp.w_row[2] = row_of(e.W13_h);
        tp.b2h = e.b12_16; tp.b3h = e.b13_16;
        const size_t row0 = (size_t)(mod_base - p.mod) / CB2_MOD_TOTAL;      // row of the current step (sampling) or of member 0 (forward)
        tp.mod16 = p.mod16 + row0 * 768 + layer * 256;
        tp.mod16_stride = mod_stride_b ? 768 : 0;
        tp.res = reinterpret_cast<const __half*>(first ? p.hE0 : p.hE);
        tp.out = reinterpret_cast<__half*>(p.hE);
    } else {
        const DecLayerW& d = m.dec[layer];
        tp.P16 = p.P16[0]; tp.n_w = 2;
        tp.w_row[0] = row_of(d.W1b2_h); tp.w_row[1] = row_of(d.W2_h); tp.b2h = d.b2_16;
    }
    const int grid = min(p.num_sms, tp.n_tiles);
#ifdef CB2_WITH_WG
    static const bool wg_impl = getenv("CB2_EDGE_WG") != nullptr;        // A/B switch of experiment builds: the warpgroup-per-tile variant
    if (wg_impl) {
        const dim3 g(grid), b(wg::WG_CTA_THREADS);
        const size_t sm = wg::wg_smem_bytes(mode);
        if (mode == EDGE_ENC_EDGE) CB2_CUDA(launch_pdl(wg::edge_wg_kernel<EDGE_ENC_EDGE, false>, g, b, sm, s, maps, tp));
        else if (mode == EDGE_ENC_NODE && masked) CB2_CUDA(launch_pdl(wg::edge_wg_kernel<EDGE_ENC_NODE, true>, g, b, sm, s, maps, tp));
        else if (mode == EDGE_ENC_NODE) CB2_CUDA(launch_pdl(wg::edge_wg_kernel<EDGE_ENC_NODE, false>, g, b, sm, s, maps, tp));
        else if (masked) CB2_CUDA(launch_pdl(wg::edge_wg_kernel<EDGE_DEC, true>, g, b, sm, s, maps, tp));
        else CB2_CUDA(launch_pdl(wg::edge_wg_kernel<EDGE_DEC, false>, g, b, sm, s, maps, tp));
        CB2_LAUNCH_CHECK();
        p.launches++;
        return 0;
    }
#endif
    const dim3 g(grid), b(CTA_THREADS);
    const size_t sm = tc_smem_bytes(mode);
    if (mode == EDGE_ENC_EDGE) CB2_CUDA(launch_pdl(edge_tc_kernel<EDGE_ENC_EDGE, false>, g, b, sm, s, maps, tp));
    else if (mode == EDGE_ENC_NODE && masked) CB2_CUDA(launch_pdl(edge_tc_kernel<EDGE_ENC_NODE, true>, g, b, sm, s, maps, tp));
    else if (mode == EDGE_ENC_NODE) CB2_CUDA(launch_pdl(edge_tc_kernel<EDGE_ENC_NODE, false>, g, b, sm, s, maps, tp));
    else if (masked) CB2_CUDA(launch_pdl(edge_tc_kernel<EDGE_DEC, true>, g, b, sm, s, maps, tp));
    else CB2_CUDA(launch_pdl(edge_tc_kernel<EDGE_DEC, false>, g, b, sm, s, maps, tp));
    CB2_LAUNCH_CHECK();
    p.launches++;
    return 0;
}

}  // namespace cb2
