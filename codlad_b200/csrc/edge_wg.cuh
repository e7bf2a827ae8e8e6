// Group-per-tile-pair version of the per-edge message MLPs (included by edge_tc.cu after TcMaps / TcParams; experiment builds:
// tools/dev/build_variant.sh NAME -DCB2_WITH_WG, selected at run time with CB2_EDGE_WG=1).
//
// Same maths as edge_f32.cu (reference models/protein_mpnn_utils.py:240-247, :261-270, :300-307), same operands and the same
// tcgen05 / TMA building blocks as edge_tc_kernel, but a different division of labour:
//
//   * one persistent 512-thread CTA per SM = TWO autonomous groups of 8 warps.  A group owns TWO tile slots (32 KB operand tile +
//     128-column TMEM accumulator each) and alternates between them stage by stage:
//         MMA1(A) MMA1(B) | E1(A) -> MMA2(A) | E1(B) -> MMA2(B) | E2(A) -> MMA3/red(A) | E2(B) -> MMA3/red(B) | E3/drain(A) | E3/drain(B)
//     so the MMA a stage hands off runs while the group works on its other slot, and a slot's next TMA load travels behind the
//     other slot's last stage.  Nothing is handed from group to group: the only synchronisation objects are the group's named
//     barrier (256 threads) and its slots' TMA / MMA-commit mbarriers.  The two groups drift out of phase, so one group's
//     operand loads and stores overlap the other's MUFU / FMA work.  (The first form of this experiment, four groups of four warps
//     with one slot each, is in the history of this file; it idled ~20 % of every tile waiting for its own MMA / TMA.)
//   * thread = (tile row, 64-column half): the TMEM lane of its warp quarter; columns in chunks of 16 / 32.  The gathered half of
//     the row (128 bytes of P16 per thread) is requested before the accumulator wait.  LayerNorm statistics of the edge update
//     are exchanged between the two halves of a row through drained TMEM columns.
//   * MMAs are issued by an elected lane of the group's warp 0 right after the group barrier that closes a stage, TMA by an
//     elected lane of its warp 1 (a thread's tcgen05.mma queues behind its own bulk copies, hence two warps).  No dedicated
//     control warps: 16 warps at up to 128 registers.
//   * waits are mbarrier.try_wait loops (the hardware suspends the warp; no nanosleep polling).
#pragma once

namespace wg {

constexpr int N_SLOT = 4;
constexpr int WG_CTA_THREADS = 512;
constexpr int N_WG_BAR = 1 + 2 * N_SLOT;     // [0] weights, [1 + s] TMA load of slot s, [5 + s] MMA commit of slot s

__device__ __forceinline__ void grp_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

// hardware-suspended wait, bounded in time so that a protocol bug traps instead of hanging the GPU
static __device__ __noinline__ void mbar_wait_hw_slow(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) mbar_timeout(bar, parity, true);
    }
}
__device__ __forceinline__ void mbar_wait_hw(uint32_t bar, uint32_t parity) {
    if (!mbar_try(bar, parity)) mbar_wait_hw_slow(bar, parity);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld2_x2(uint32_t taddr, uint32_t col_stride, float (&v)[4]) {      // 2 x (2 columns), one wait
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[2]), "=r"(r[3]) : "r"(taddr + col_stride));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// 64 bytes (32 halves) from global memory through the read-only path
__device__ __forceinline__ void ldg64B(const __half* src, uint32_t (&v)[16]) {
    ldg256(src, *reinterpret_cast<uint32_t(*)[8]>(&v[0]));
    ldg256(src + 16, *reinterpret_cast<uint32_t(*)[8]>(&v[8]));
}

template <int MODE, bool MASKED>
__global__ void __launch_bounds__(WG_CTA_THREADS, 1) edge_wg_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int N_W = MODE == EDGE_ENC_EDGE ? 3 : 2;
    // layout: [weights N_W x 32 KB][4 operand tiles x 32 KB][indicator 4 KB (ENC_NODE / DEC)][barriers][TMEM base]
    unsigned char* sW = smem;
    unsigned char* sT = sW + N_W * TILE_BYTES;
    unsigned char* sAux = sT + N_SLOT * TILE_BYTES;
    unsigned char* sInd = sAux;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sAux + (MODE == EDGE_ENC_EDGE ? 0 : IND_BYTES));
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + N_WG_BAR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 3, wq = warp & 3, ch = (warp >> 2) & 1, r = wq * 32 + lane;      // group, TMEM lane quarter, column half, tile row
    const int K = p.K, NPT = p.NPT;
    if (tid == 0) trace_window(p.trace, p.trace_slot, false);
    if ((smem_u32(smem) & 1023u) != 0) __trap();                                // SWIZZLE_128B atoms repeat every 1 KiB
    pdl_launch_dependents();               // the next kernel of the step may start its prologue on SMs this grid leaves idle

    // ---- one-time setup: barriers, TMEM, indicator operand of the reduction MMA ----
    if (tid == 0) {
        for (int i = 0; i < N_WG_BAR; ++i) mbar_init(smem_u32(&sBar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (MODE != EDGE_ENC_EDGE) {
        // Ind[q][r] = 1/2 if row r belongs to node q (K-major SW128, 16 rows x 128 k); 1/2: the activations are 2 GELU
        for (int t = tid; t < 16 * 16; t += WG_CTA_THREADS) {
            const int q = t >> 4, c16 = t & 15;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r0 = c16 * 8 + e * 2;
                const __half lo = __float2half((q < NPT && r0 / K == q) ? 0.5f : 0.f);
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
    const uint32_t bar_w = smem_u32(&sBar[0]);
    const bool mma_warp = (warp & 7) == 0, tma_warp = (warp & 7) == 1;
    if (warp == 1) {                                                            // the layer weights: static data, fetched before the dependency wait
        if (elect_one()) {
            mbar_expect_tx(bar_w, (uint32_t)(N_W * TILE_BYTES));
            for (int m = 0; m < N_W; ++m)
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(smem_u32(sW + m * TILE_BYTES + h * HALF_BYTES), &maps.weights, h * 64, p.w_row[m], bar_w);
        }
        __syncwarp();
    }
    // contiguous, balanced tile range of this CTA (sizes differ by at most one tile); slot s takes tiles tile_begin + s + 4 i
    const int tile_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
    const int tile_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
    constexpr uint32_t IDESC_MAIN = umma_idesc(128, 128, 0, 0);
    constexpr uint32_t IDESC_RED = umma_idesc(128, 16, 1, 0);
    const CUtensorMap* in_map = p.in_is_frame ? &maps.in_frame : &maps.state;
    const int q_of_r = r / K, k_of_r = r - q_of_r * K;
    const int len0 = __ldg(p.lengths);

    // ---- per-slot addresses (q = 0 / 1: the group's first / second slot) ----
    auto slot_T = [&](int q) { return sT + (2 * grp + q) * TILE_BYTES; };
    auto slot_bar_load = [&](int q) { return smem_u32(&sBar[1 + 2 * grp + q]); };
    auto slot_bar_acc = [&](int q) { return smem_u32(&sBar[1 + N_SLOT + 2 * grp + q]); };
    auto slot_tmem = [&](int q) { return tmem_base + (uint32_t)((2 * grp + q) * 128); };
    auto slot_tmem_row = [&](int q) { return slot_tmem(q) + ((uint32_t)(wq * 32) << 16); };

    // ---- tile bookkeeping (all warp-uniform except the row's neighbour index) ----
    struct Tile { int b, i0, nv, in_row0, f; };
    auto tile_of = [&](int t) {
        Tile x;
        x.b = t / p.tiles_per_member;
        x.i0 = (t - x.b * p.tiles_per_member) * NPT;
        x.nv = min(NPT, p.L - x.i0);
        x.f = p.single_frame ? 0 : __ldg(p.frame_of + x.b);
        x.in_row0 = ((p.in_is_frame ? x.f : x.b) * p.L + x.i0) * K;
        return x;
    };
    auto load_j = [&](int t) -> int {                  // member-local index of this row's neighbour in tile t, or -1 for a row outside it
        if (t >= tile_end) return -1;
        const Tile x = tile_of(t);
        if (q_of_r >= x.nv) return -1;
        return __ldg(p.nbr_idx + ((size_t)x.f * p.L + x.i0 + q_of_r) * K + k_of_r);
    };
    auto issue_load = [&](int q, int t) {              // TMA of tile t's h_E rows into slot q (tma_warp only)
        const Tile x = tile_of(t);
        const uint32_t T_u32 = smem_u32(slot_T(q)), bar = slot_bar_load(q);
        if (elect_one()) {
            mbar_expect_tx(bar, (uint32_t)(x.nv * K * 256));
            for (int n = 0; n < x.nv; ++n)
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(T_u32 + h * HALF_BYTES + n * K * 128, in_map, h * 64, x.in_row0 + n * K, bar);
        }
        __syncwarp();
    };
    auto issue_mma = [&](int q, int w_slot) {          // 128x128x128 GEMM: A = slot q's tile, B = weight block w_slot (mma_warp only)
        tc_fence_after();
        const uint64_t a0 = umma_desc(smem_u32(slot_T(q)), 16, 1024), b0 = umma_desc(smem_u32(sW + w_slot * TILE_BYTES), 16, 1024);
        const uint32_t d = slot_tmem(q), bar = slot_bar_acc(q);
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint64_t koff = (uint64_t)(((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4);
                umma_f16(d, a0 + koff, b0 + koff, IDESC_MAIN, k > 0);
            }
            umma_commit(bar);
        }
        __syncwarp();
    };
    auto issue_reduce = [&](int q) {                   // D[c, n] = sum_r G[r, c] Ind[n, r]: the activation tile as an MN-major A operand
        tc_fence_after();
        const uint64_t a0 = umma_desc(smem_u32(slot_T(q)), HALF_BYTES, 1024), b0 = umma_desc(smem_u32(sInd), 16, 1024);
        const uint32_t d = slot_tmem(q), bar = slot_bar_acc(q);
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_f16(d, a0 + (uint64_t)((k * 16 * 128) >> 4), b0 + (uint64_t)(((k >> 2) * (16 * 128) + (k & 3) * 32) >> 4), IDESC_RED, k > 0);
            umma_commit(bar);
        }
        __syncwarp();
    };
    // a stage is closed by: my shared-memory writes visible to the async proxy, my TMEM reads ordered, the group converged
    auto close_stage = [&](bool wrote_smem) {
        if (wrote_smem) fence_async_smem();
        tc_fence_before();
        grp_sync(grp);
    };

    pdl_wait();                                        // h_E / P16 / S belong to the previous kernels of the step
    const int t_first = tile_begin + 2 * grp;          // tiles of this group's slots in iteration i: t_first + 4 i (+ 1)
    if (t_first < tile_end) {
        if (tma_warp) {
            issue_load(0, t_first);
            if (t_first + 1 < tile_end) issue_load(1, t_first + 1);
        }
        int jA = load_j(t_first), jB = load_j(t_first + 1);
        if (mma_warp) mbar_wait_hw(bar_w, 0);          // weights resident
        for (int it = 0, tA = t_first; tA < tile_end; ++it, tA += N_SLOT) {
            const int n_live = (tA + 1 < tile_end) ? 2 : 1;                       // slot B has a tile in this iteration?
            const int jnA = load_j(tA + N_SLOT), jnB = load_j(tA + 1 + N_SLOT);   // consumed a whole iteration later
            const uint32_t ph_load = (uint32_t)(it & 1);
            const uint32_t ph1 = (uint32_t)((3 * it) & 1), ph2 = ph1 ^ 1u, ph3 = ph1;          // three accumulator phases per tile
            // ---------------- MMA 1 of both slots: W?b . h_E ----------------
            if (mma_warp)
                for (int q = 0; q < n_live; ++q) { mbar_wait_hw(slot_bar_load(q), ph_load); issue_mma(q, 0); }
            // ---------------- E1: 2 GELU(acc + Pa[i] + Pc[j]) -> fp16 activation tile, then MMA 2 ----------------
#pragma unroll 1
            for (int q = 0; q < n_live; ++q) {
                const Tile cur = tile_of(tA + q);
                const int j = q ? jB : jA;
                const bool row_in_tile = j >= 0;
                const int node0 = cur.b * p.L + cur.i0;
                const __half* pc_src = p.P16 + ((size_t)cur.b * p.L + (row_in_tile ? j : 0)) * 256 + 128 + ch * 64;     // gathered half, my 64 columns
                const __half* pa_src = p.P16 + (size_t)(node0 + (row_in_tile ? q_of_r : 0)) * 256 + ch * 64;            // own half
                unsigned char* T = slot_T(q);
                uint32_t pc0[8], pc1[8], pc2[8], pc3[8];
                ldg256(pc_src, pc0); ldg256(pc_src + 16, pc1); ldg256(pc_src + 32, pc2); ldg256(pc_src + 48, pc3);
                mbar_wait_hw(slot_bar_acc(q), ph1);
                tc_fence_after();
                const uint32_t trow = slot_tmem_row(q) + (uint32_t)(ch * 64);
                auto chunk = [&](int c, const uint32_t (&pc)[8]) {
                    uint32_t pa[8];
                    ldg256(pa_src + c * 16, pa);                              // L1 broadcast: the rows of a node share it
                    float acc[16];
                    tmem_ld16(trow + (uint32_t)(c * 16), acc);
                    uint32_t o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const __half2 x = __hadd2(__hadd2(as_h2(pack_sat(acc[2 * e], acc[2 * e + 1])), as_h2(pa[e])), as_h2(pc[e]));
                        o[e] = as_u32(gelu2_h2(x));
                    }
                    *reinterpret_cast<uint4*>(T + tile_off(r, ch * 8 + c * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(T + tile_off(r, ch * 8 + c * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
                };
                chunk(0, pc0); chunk(1, pc1); chunk(2, pc2); chunk(3, pc3);
                close_stage(true);
                if (mma_warp) issue_mma(q, 1);
            }
            // ---------------- E2: 2 GELU(acc + b2) (rows outside the neighbour sum -> 0), then the reduction MMA / MMA 3 ----------------
#pragma unroll 1
            for (int q = 0; q < n_live; ++q) {
                unsigned char* T = slot_T(q);
                uint32_t keep = 0xffffffffu;
                if (MODE != EDGE_ENC_EDGE && MASKED) {
                    // reference: mask_attend = mask_i mask_j in the encoder; the decoder passes no mask (only rows outside the tile drop)
                    const int j = q ? jB : jA;
                    keep = j >= 0 ? 0xffffffffu : 0u;
                    if (MODE == EDGE_ENC_NODE && j >= 0) {
                        const Tile cur = tile_of(tA + q);
                        const int len = p.single_frame ? len0 : __ldg(p.lengths + cur.f);
                        keep = (cur.i0 + q_of_r < len && j < len) ? 0xffffffffu : 0u;
                    }
                }
                mbar_wait_hw(slot_bar_acc(q), ph2);
                tc_fence_after();
                const uint32_t trow = slot_tmem_row(q) + (uint32_t)(ch * 64);
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t bb[16];
                    ldg64B(p.b2h + ch * 64 + c * 32, bb);
                    float acc[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), acc);
                    uint32_t o[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const __half2 x = __hadd2(as_h2(pack_sat(acc[2 * e], acc[2 * e + 1])), as_h2(bb[e]));
                        o[e] = as_u32(gelu2_h2(x));
                        if (MODE != EDGE_ENC_EDGE && MASKED) o[e] &= keep;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<uint4*>(T + tile_off(r, ch * 8 + c * 4 + k)) = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
                }
                close_stage(true);
                if (mma_warp) { if (MODE == EDGE_ENC_EDGE) issue_mma(q, 2); else issue_reduce(q); }
            }
            // ---------------- third stage: drain of the neighbour sums, or E3 of the edge update ----------------
#pragma unroll 1
            for (int q = 0; q < n_live; ++q) {
                const Tile cur = tile_of(tA + q);
                const int t_next = tA + q + N_SLOT;
                const bool row_in_tile = (q ? jB : jA) >= 0;
                unsigned char* T = slot_T(q);
                if (MODE != EDGE_ENC_EDGE) {
                    // S[node, c]: lane = output column c (128 lanes = the ch == 0 half of the group), columns = the tile's nodes
                    mbar_wait_hw(slot_bar_acc(q), ph3);
                    tc_fence_after();
                    if (tma_warp && t_next < tile_end) issue_load(q, t_next);          // the operand tile is free: fetch the next one behind the drain
                    if (ch == 0) {
                        float s4[4];
                        tmem_ld4(slot_tmem_row(q), s4);
                        const int node0 = cur.b * p.L + cur.i0;
#pragma unroll
                        for (int n = 0; n < MAX_NPT; ++n)
                            if (n < cur.nv) p.S[((size_t)node0 + n) * 128 + r] = s4[n];
                    }
                    close_stage(false);                                          // accumulator drained by every thread before the next MMA 1
                } else {
                    // + residual, LayerNorm, adaLN modulate / gate -> fp16 tile -> TMA store
                    const __half* res_src = p.res + ((size_t)cur.in_row0 + (row_in_tile ? r : 0)) * 128 + ch * 64;      // the row's own h_E (L2)
                    uint32_t rs0[8], rs1[8], rs2[8], rs3[8];
                    ldg256_coherent(res_src, rs0); ldg256_coherent(res_src + 16, rs1); ldg256_coherent(res_src + 32, rs2); ldg256_coherent(res_src + 48, rs3);
                    mbar_wait_hw(slot_bar_acc(q), ph3);
                    tc_fence_after();
                    const uint32_t trow = slot_tmem_row(q) + (uint32_t)(ch * 64);
                    // pass A: v = residual + (acc + b13) in packed half, parked in the tile (in place: this thread's own row and columns);
                    // statistics of those values in fp32
                    float sum = 0.f, sq = 0.f;
                    auto chunk_a = [&](int c, const uint32_t (&rs)[8]) {
                        uint32_t b3[8];
                        ldg256(p.b3h + ch * 64 + c * 16, b3);
                        float acc[16];
                        tmem_ld16(trow + (uint32_t)(c * 16), acc);
                        uint32_t o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const __half2 v = __hadd2(as_h2(rs[e]), __hadd2(as_h2(pack_sat(acc[2 * e], acc[2 * e + 1])), as_h2(b3[e])));
                            const float2 vf = __half22float2(v);
                            sum += vf.x + vf.y;
                            sq = fmaf(vf.x, vf.x, fmaf(vf.y, vf.y, sq));
                            o[e] = as_u32(v);
                        }
                        *reinterpret_cast<uint4*>(T + tile_off(r, ch * 8 + c * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(T + tile_off(r, ch * 8 + c * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
                    };
                    chunk_a(0, rs0); chunk_a(1, rs1); chunk_a(2, rs2); chunk_a(3, rs3);
                    // the two column halves of a row meet in accumulator columns each has drained itself (its own first two): tcgen05.st,
                    // group barrier, tcgen05.ld; summed in a fixed order (half 0 + half 1) -> deterministic
                    tmem_st2(trow, sum, sq);
                    tc_fence_before();
                    grp_sync(grp);
                    tc_fence_after();
                    float st4[4];
                    tmem_ld2_x2(slot_tmem_row(q), 64u, st4);
                    const float tsum = st4[0] + st4[2], tsq = st4[1] + st4[3];
                    const float mean = tsum * (1.0f / 128.0f);
                    const float rstd = rsqrtf(fmaxf(tsq * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-6f);
                    const __half2 rstd2 = __float2half2_rn(rstd), nmr2 = __float2half2_rn(-mean * rstd);
                    const __half* mod_row = p.mod16 + (size_t)cur.b * p.mod16_stride + ch * 64;
                    // pass B (packed half): out = (v rstd - mean rstd) A[c] + B[c],  A = gate (1 + scale), B = gate * shift
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t av[16], bv[16], o[16];
                        ldg64B(mod_row + c * 32, av);
                        ldg64B(mod_row + 128 + c * 32, bv);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint4 v4 = *reinterpret_cast<const uint4*>(T + tile_off(r, ch * 8 + c * 4 + k));
                            const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                o[4 * k + e] = as_u32(__hfma2(__hfma2(as_h2(vv[e]), rstd2, nmr2), as_h2(av[4 * k + e]), as_h2(bv[4 * k + e])));
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            *reinterpret_cast<uint4*>(T + tile_off(r, ch * 8 + c * 4 + k)) = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
                    }
                    close_stage(true);
                    if (tma_warp) {
                        const int out_row0 = (cur.b * p.L + cur.i0) * K;
                        const uint32_t T_u32 = smem_u32(T);
                        if (elect_one()) {                // (bulk-copy groups belong to the issuing thread: elect.sync picks the same lane every time)
                            for (int n = 0; n < cur.nv; ++n)
                                for (int h = 0; h < 2; ++h)
                                    tma_store_2d(&maps.state, h * 64, out_row0 + n * K, T_u32 + h * HALF_BYTES + n * K * 128);
                            tma_store_commit();
                            tma_store_wait_read();        // the store has finished reading the tile: it may be overwritten
                        }
                        __syncwarp();
                        if (t_next < tile_end) issue_load(q, t_next);
                    }
                }
            }
            jA = jnA;
            jB = jnB;
        }
        if (MODE == EDGE_ENC_EDGE && tma_warp) { if (elect_one()) tma_store_wait_all(); __syncwarp(); }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) trace_window(p.trace, p.trace_slot, true);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

inline size_t wg_smem_bytes(int mode) {
    const int n_w = mode == EDGE_ENC_EDGE ? 3 : 2;
    return (size_t)(n_w + N_SLOT) * TILE_BYTES + (mode == EDGE_ENC_EDGE ? 0 : IND_BYTES) + N_WG_BAR * 8 + 16;
}

}  // namespace wg
