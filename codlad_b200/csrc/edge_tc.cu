// tcgen05 tier of the per-edge message MLPs (the dominant kernels of the denoiser step).
//
// Same maths as edge_f32.cu (reference models/protein_mpnn_utils.py:240-247, :261-270, :300-307), but the
// chained 128x128 linear layers run on the 5th-generation tensor cores:
//
//   * edge state h_E, activations and weights are fp16 (kind::f16 MMA, fp32 accumulation in TMEM).  fp16 rather
//     than bf16: same tensor rate, 8x finer mantissa, and every operand here is LayerNorm/GELU-bounded.
//   * one CTA per SM (persistent), NWG independent 128-thread "tile pipelines" per CTA that share the layer
//     weights resident in shared memory (SWIZZLE_128B, K-major, loaded once by TMA).  A tile = NPT whole nodes
//     (NPT*K <= 128 neighbour rows) of one ensemble member, so the neighbour reduction is tile-local.
//   * per tile: TMA loads the h_E rows (contiguous in HBM) into a swizzled K-major A tile -> MMA 1 (W?b h_E) ->
//     epilogue 1 reads the accumulator from TMEM (thread = row), adds the per-node halves of the first layer
//     (Pa[i] broadcast from smem, Pc[j] gathered from L2 as fp16 with 256-bit loads), GELU, writes the fp16
//     activation back into the same smem tile -> MMA 2 -> epilogue 2 (+b, GELU) ->
//        ENC_NODE / DEC : the masked neighbour sum is one more (tiny) MMA: S^T[c, q] = sum_r G[r, c] * Ind[q, r]
//                         with the activation tile reused as an MN-major A operand and a 16-row indicator B;
//        ENC_EDGE       : MMA 3 (W13) -> epilogue 3: +residual, LayerNorm, adaLN modulate/gate -> fp16 tile ->
//                         TMA store.
//     While one pipeline is in an epilogue the tensor core runs another pipeline's MMAs.
//   * GELU uses tanh.approx (one MUFU per element) on a refitted 2-term inner polynomial: |err| < 2.8e-4 abs
//     against erf-GELU before the MUFU's own 2^-11 relative error -- the same order as the fp16 rounding of the
//     activation it feeds.  (The fp32 tier uses erff.)
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
    int mode, L, K, NPT, tiles_per_member, n_tiles;
    int in_is_frame;                 // layer 0 of the encoder reads h_E0 (indexed by frame)
    int w_row[3];                    // first row of each weight block in the packed weight tensor
    int n_w;                         // 2 (ENC_NODE / DEC) or 3 (ENC_EDGE)
    const float* P;                  // [N, 256]: [:, :128] = Wa h_V_i + b1 (own half)
    const __half* Pc;                // [N, 128]: gathered half Wc h_V_j (+ decoder table), fp16
    const float *b2, *b3;            // second / third layer biases (fp32)
    const float* mod;                // ENC_EDGE: adaLN block of this layer; member row at mod + b * mod_stride
    int mod_stride;
    const __half* res;               // ENC_EDGE: residual source (= the tile's input rows)
    const int *lengths, *frame_of, *nbr_idx;
    float* S;                        // [N, 128] neighbour sums (ENC_NODE / DEC)
};

// ---------------------------------------------------------------------------------------------- the kernel
// packed-half helpers of the epilogues
__device__ __forceinline__ uint32_t pack_sat(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// erf-GELU ~= 0.5 x (1 + tanh(x (a + b x^2))) on two fp16 values (coefficients refitted, see gelu_fast)
__device__ __forceinline__ __half2 gelu_h2(__half2 x) {
    const __half2 x2 = __hmul2(x, x);
    const __half2 pl = __hfma2(x2, __float2half2_rn(0.03470094f), __float2half2_rn(0.80015698f));
    const __half2 u = __hmul2(x, pl);
    uint32_t t;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(as_u32(u)));
    const __half2 h = __hmul2(x, __float2half2_rn(0.5f));
    return __hfma2(h, as_h2(t), h);
}

struct TileMeta {
    int b, i0, nv;
    int in_row0, out_row0;
    size_t node0;
    bool row_valid, row_on;
    int q_of_r;
    const __half* pc_row;
    float pa[MAX_NPT];
    float modA, modB;        // ENC_EDGE: gate (1 + scale) and gate * shift of this thread's column
};

template <int NWG, int MODE>
__global__ void __launch_bounds__(NWG * 128, 1) edge_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B atoms repeat every 1 KiB
    // layout: [weights n_w x 32 KB][tiles NWG x 32 KB][indicator 4 KB][CTA vectors][per-WG vectors][barriers]
    unsigned char* sW = smem;
    unsigned char* sT = sW + p.n_w * TILE_BYTES;
    unsigned char* sInd = sT + NWG * TILE_BYTES;
    __half* sB2h = reinterpret_cast<__half*>(sInd + IND_BYTES);          // [128] second-layer bias, fp16
    float* sB3 = reinterpret_cast<float*>(sB2h + 128);                   // [128] third-layer bias, fp32
    __half* sWgVec = reinterpret_cast<__half*>(sB3 + 128);               // per WG: Pa[MAX_NPT][128], modA[128], modB[128]
    constexpr int VEC_PER_WG = (MAX_NPT + 2) * 128;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sWgVec + NWG * VEC_PER_WG);   // [0] weights, [1 + 2 wg] load, [2 + 2 wg] mma
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 1 + 2 * NWG);

    const int tid = threadIdx.x, wg = tid >> 7, wt = tid & 127, warp_in_wg = wt >> 5;
    const int K = p.K, NPT = p.NPT;

    // ---- one-time setup: barriers, TMEM, weights, indicator operand ----
    if (tid == 0) {
        mbar_init(smem_u32(&sBar[0]), 1);
        for (int g = 0; g < NWG; ++g) { mbar_init(smem_u32(&sBar[1 + 2 * g]), 1); mbar_init(smem_u32(&sBar[2 + 2 * g]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // indicator B operand of the reduction MMA: Ind[q][r] = 1 if row r belongs to node q (K-major SW128, 16 rows x 128 k)
    for (int t = tid; t < 16 * 16; t += NWG * 128) {
        const int q = t >> 4, c16 = t & 15;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r0 = c16 * 8 + e * 2;
            const __half lo = __float2half((q < NPT && r0 / K == q) ? 1.f : 0.f);
            const __half hi = __float2half((q < NPT && (r0 + 1) / K == q) ? 1.f : 0.f);
            w[e] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
        }
        const uint32_t off = (uint32_t)((c16 >> 3) * (16 * 128) + q * 128 + (((c16 & 7) ^ (q & 7)) << 4));
        *reinterpret_cast<uint4*>(sInd + off) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (tid < 128) {
        sB2h[tid] = __float2half_rn(p.b2[tid]);
        sB3[tid] = MODE == EDGE_ENC_EDGE ? p.b3[tid] : 0.f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    if (tid == 0) {
        mbar_expect_tx(smem_u32(&sBar[0]), (uint32_t)(p.n_w * TILE_BYTES));
        for (int m = 0; m < p.n_w; ++m)
            for (int h = 0; h < 2; ++h)
                tma_load_2d(smem_u32(sW + m * TILE_BYTES + h * HALF_BYTES), &maps.weights, h * 64, p.w_row[m], smem_u32(&sBar[0]));
    }

    unsigned char* T = sT + wg * TILE_BYTES;
    const uint32_t T_u32 = smem_u32(T);
    __half* sPa = sWgVec + wg * VEC_PER_WG;
    __half* sModA = sPa + MAX_NPT * 128;
    __half* sModB = sModA + 128;
    const uint32_t bar_load = smem_u32(&sBar[1 + 2 * wg]), bar_mma = smem_u32(&sBar[2 + 2 * wg]);
    const uint32_t tmem_acc = tmem_base + (uint32_t)(wg * 128);                       // this pipeline's 128 accumulator columns
    const uint32_t tmem_row = tmem_acc + ((uint32_t)(warp_in_wg * 32) << 16);         // + this warp's lane quarter
    const CUtensorMap* in_map = p.in_is_frame ? &maps.in_frame : &maps.state;
    constexpr uint32_t IDESC_MAIN = umma_idesc(128, 128, 0, 0);
    constexpr uint32_t IDESC_RED = umma_idesc(128, 16, 1, 0);
    uint32_t ph_load = 0, ph_mma = 0;
    bool weights_ready = false;
    const int r = wt;                          // this thread's tile row == TMEM lane
    const int tile_stride = gridDim.x * NWG;

    // per-tile metadata: every global load it needs is issued one tile ahead
    auto load_meta = [&](int tile) {
        TileMeta m;
        m.b = tile / p.tiles_per_member;
        m.i0 = (tile - m.b * p.tiles_per_member) * NPT;
        m.nv = min(NPT, p.L - m.i0);
        const int f = p.frame_of[m.b];
        const int len = p.lengths[f];
        m.node0 = (size_t)m.b * p.L + m.i0;
        m.in_row0 = (int)(((size_t)(p.in_is_frame ? f : m.b) * p.L + m.i0) * K);
        m.out_row0 = (int)(m.node0 * K);
        m.q_of_r = r / K;
        m.row_valid = m.q_of_r < m.nv;
        int j = 0;
        m.row_on = false;
        if (m.row_valid) {
            const int i = m.i0 + m.q_of_r;
            j = p.nbr_idx[((size_t)f * p.L + i) * K + (r - m.q_of_r * K)];
            m.row_on = (MODE == EDGE_DEC) ? true : (i < len && j < len);
        }
        m.pc_row = p.Pc + ((size_t)m.b * p.L + j) * 128;
#pragma unroll
        for (int q = 0; q < MAX_NPT; ++q) m.pa[q] = q < m.nv ? p.P[(m.node0 + q) * 256 + wt] : 0.f;
        m.modA = 0.f; m.modB = 0.f;
        if (MODE == EDGE_ENC_EDGE) {
            const float* md = p.mod + (size_t)m.b * p.mod_stride;
            const float gate = md[1024 + wt];
            m.modA = gate * (1.0f + md[896 + wt]);
            m.modB = gate * md[768 + wt];
        }
        return m;
    };
    auto issue_load = [&](const TileMeta& m) {          // one thread: TMA of the tile's h_E rows, one box per node and half
        mbar_expect_tx(bar_load, (uint32_t)(m.nv * K * 256));
        for (int q = 0; q < m.nv; ++q)
            for (int h = 0; h < 2; ++h)
                tma_load_2d(T_u32 + h * HALF_BYTES + q * K * 128, in_map, h * 64, m.in_row0 + q * K, bar_load);
    };
    auto issue_mma = [&](int w_slot) {                  // one thread: 128x128x128 GEMM, A = activation tile, B = weight slot
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t koff = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);
            umma_f16(tmem_acc, umma_desc(T_u32 + koff, 16, 1024), umma_desc(smem_u32(sW + w_slot * TILE_BYTES) + koff, 16, 1024), IDESC_MAIN, k > 0);
        }
        umma_commit(bar_mma);
    };

    int tile = blockIdx.x * NWG + wg;
    TileMeta nxt{};
    if (tile < p.n_tiles) {
        nxt = load_meta(tile);
        if (wt == 0) issue_load(nxt);
    }
    while (tile < p.n_tiles) {
        const TileMeta cur = nxt;
        const int next_tile = tile + tile_stride;
        // ---- stage this tile's per-node / per-member vectors ----
#pragma unroll
        for (int q = 0; q < MAX_NPT; ++q) sPa[q * 128 + wt] = __float2half_rn(cur.pa[q]);
        if (MODE == EDGE_ENC_EDGE) { sModA[wt] = __float2half_rn(cur.modA); sModB[wt] = __float2half_rn(cur.modB); }
        wg_sync(wg);
        if (!weights_ready) { mbar_wait(smem_u32(&sBar[0]), 0); weights_ready = true; }
        uint32_t pc0[16], pc1[16];                       // gathered Pc[j] chunks (32 halves each), double buffered
        {
            uint32_t (&lo)[8] = *reinterpret_cast<uint32_t(*)[8]>(&pc0[0]);
            uint32_t (&hi)[8] = *reinterpret_cast<uint32_t(*)[8]>(&pc0[8]);
            ldg256(cur.pc_row, lo); ldg256(cur.pc_row + 16, hi);
        }
        mbar_wait(bar_load, ph_load); ph_load ^= 1;

        // ---- MMA 1: acc = h_E . W?b^T ; then prefetch the next tile's metadata while it runs ----
        if (wt == 0) issue_mma(0);
        if (next_tile < p.n_tiles) nxt = load_meta(next_tile);
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();

        // ---- epilogue 1: GELU(acc + Pa[i] + Pc[j]) -> fp16 activation tile (packed-half arithmetic) ----
        {
            const __half* pa_row = sPa + (cur.row_valid ? cur.q_of_r : 0) * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t (&pcc)[16] = (c & 1) ? pc1 : pc0;
                uint32_t (&pcn)[16] = (c & 1) ? pc0 : pc1;
                if (c < 3) {
                    uint32_t (&lo)[8] = *reinterpret_cast<uint32_t(*)[8]>(&pcn[0]);
                    uint32_t (&hi)[8] = *reinterpret_cast<uint32_t(*)[8]>(&pcn[8]);
                    ldg256(cur.pc_row + (c + 1) * 32, lo); ldg256(cur.pc_row + (c + 1) * 32 + 16, hi);
                }
                float acc[32];
                tmem_ld32(tmem_row + c * 32, acc);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint4 pav = *reinterpret_cast<const uint4*>(pa_row + c * 32 + u * 8);
                    const uint32_t pa4[4] = {pav.x, pav.y, pav.z, pav.w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 x = __hadd2(__hadd2(as_h2(pack_sat(acc[u * 8 + e * 2], acc[u * 8 + e * 2 + 1])), as_h2(pa4[e])), as_h2(pcc[u * 4 + e]));
                        o[e] = as_u32(gelu_h2(x));
                    }
                    *reinterpret_cast<uint4*>(T + tile_off(r, c * 4 + u)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        wg_sync(wg);

        // ---- MMA 2 ----
        if (wt == 0) issue_mma(1);
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();

        // ---- epilogue 2: GELU(acc + b2) (masked rows -> 0 for the reduction) -> fp16 tile ----
        {
            const uint32_t keep = (MODE == EDGE_ENC_EDGE || cur.row_on) ? 0xffffffffu : 0u;     // bit mask: no branch per pair
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float acc[32];
                tmem_ld32(tmem_row + c * 32, acc);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint4 bv = *reinterpret_cast<const uint4*>(sB2h + c * 32 + u * 8);
                    const uint32_t b4[4] = {bv.x, bv.y, bv.z, bv.w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 x = __hadd2(as_h2(pack_sat(acc[u * 8 + e * 2], acc[u * 8 + e * 2 + 1])), as_h2(b4[e]));
                        o[e] = as_u32(gelu_h2(x)) & keep;
                    }
                    *reinterpret_cast<uint4*>(T + tile_off(r, c * 4 + u)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        wg_sync(wg);

        if (MODE != EDGE_ENC_EDGE) {
            // ---- neighbour sum as an MMA: D[c, q] = sum_r G[r, c] Ind[q, r]  (A = G^T, MN-major view of the tile) ----
            if (wt == 0) {
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t a = umma_desc(T_u32 + (uint32_t)(k * 16 * 128), HALF_BYTES, 1024);      // 16 rows (K) per step
                    const uint64_t bd = umma_desc(smem_u32(sInd) + (uint32_t)((k >> 2) * (16 * 128) + (k & 3) * 32), 16, 1024);
                    umma_f16(tmem_acc, a, bd, IDESC_RED, k > 0);
                }
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
            tc_fence_after();
            if (wt == 0 && next_tile < p.n_tiles) issue_load(nxt);       // T is free: start the next tile's TMA now
            float s4[4];
            tmem_ld4(tmem_row, s4);             // lane = output column, 4 columns = nodes of the tile
#pragma unroll
            for (int q = 0; q < MAX_NPT; ++q)
                if (q < cur.nv) p.S[(cur.node0 + q) * 128 + wt] = s4[q];
            tc_fence_before();
            wg_sync(wg);                        // TMEM and the per-tile vectors are reused by the next tile
        } else {
            // ---- MMA 3 (W13), then residual + LayerNorm + adaLN -> fp16 tile -> TMA store ----
            const __half* res_row = p.res + ((size_t)cur.in_row0 + (cur.row_valid ? r : 0)) * 128;
            uint32_t rs0[16], rs1[16];
            {
                uint32_t (&lo)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rs0[0]);
                uint32_t (&hi)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rs0[8]);
                ldg256_coherent(res_row, lo); ldg256_coherent(res_row + 16, hi);
            }
            if (wt == 0) issue_mma(2);
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
            tc_fence_after();
            // pass A (fp32): v = residual + acc + b13, row statistics; v is parked as fp16 in this thread's tile row
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t (&rc)[16] = (c & 1) ? rs1 : rs0;
                uint32_t (&rn)[16] = (c & 1) ? rs0 : rs1;
                if (c < 3) {
                    uint32_t (&lo)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rn[0]);
                    uint32_t (&hi)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rn[8]);
                    ldg256_coherent(res_row + (c + 1) * 32, lo); ldg256_coherent(res_row + (c + 1) * 32 + 16, hi);
                }
                float acc[32];
                tmem_ld32(tmem_row + c * 32, acc);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = u * 8 + e * 2;
                        const float2 rr = h2_to_f2(rc[u * 4 + e]);
                        const float2 bb = *reinterpret_cast<const float2*>(sB3 + c * 32 + col);
                        const float v0 = rr.x + (acc[col] + bb.x), v1 = rr.y + (acc[col + 1] + bb.y);
                        sum += v0 + v1;
                        sq = fmaf(v0, v0, fmaf(v1, v1, sq));
                        o[e] = pack_sat(v0, v1);
                    }
                    *reinterpret_cast<uint4*>(T + tile_off(r, c * 4 + u)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            const float mean = sum * (1.0f / 128.0f);
            const float var = fmaxf(sq * (1.0f / 128.0f) - mean * mean, 0.f);
            const __half2 rstd2 = __float2half2_rn(rsqrtf(var + 1e-6f));
            const __half2 mean2 = __float2half2_rn(mean);
            // pass B (packed half): out = (v - mean) * (rstd * A[c]) + B[c],  A = gate (1 + scale), B = gate * shift
#pragma unroll
            for (int c16 = 0; c16 < 16; ++c16) {
                uint4* slot = reinterpret_cast<uint4*>(T + tile_off(r, c16));
                const uint4 vv = *slot;
                const uint4 av = *reinterpret_cast<const uint4*>(sModA + c16 * 8);
                const uint4 bv = *reinterpret_cast<const uint4*>(sModB + c16 * 8);
                const uint32_t v4[4] = {vv.x, vv.y, vv.z, vv.w}, a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    o[e] = as_u32(__hfma2(__hsub2(as_h2(v4[e]), mean2), __hmul2(rstd2, as_h2(a4[e])), as_h2(b4[e])));
                *slot = make_uint4(o[0], o[1], o[2], o[3]);
            }
            fence_async_smem();
            tc_fence_before();
            wg_sync(wg);
            if (wt == 0) {
                for (int q = 0; q < cur.nv; ++q)
                    for (int h = 0; h < 2; ++h)
                        tma_store_2d(&maps.state, h * 64, cur.out_row0 + q * K, T_u32 + h * HALF_BYTES + q * K * 128);
                tma_store_commit();
                if (next_tile < p.n_tiles) {
                    tma_store_wait_read();                               // the store has finished reading T
                    issue_load(nxt);
                }
            }
        }
        tile = next_tile;
    }
    if (MODE == EDGE_ENC_EDGE && wt == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

size_t tc_smem_bytes(int nwg, int n_w) {
    return (size_t)n_w * TILE_BYTES + (size_t)nwg * TILE_BYTES + IND_BYTES + 128 * 2 + 128 * 4 + (size_t)nwg * (MAX_NPT + 2) * 128 * 2 +
           (1 + 2 * nwg) * 8 + 16 + 1024 /* alignment slack */;
}

}  // namespace

int edge_tc_prepare(Plan& p) {
    if (p.K > 128 || p.K < 1) { set_error("edge_tc: K=%d unsupported", p.K); return 1; }
    EncodeTiledFn fn = nullptr;
    if (get_encode_fn(&fn)) return 1;
    TcMaps* m = new TcMaps();
    p.tmaps = m;
    if (encode_rows_map(fn, &m->in_frame, p.hE0, (size_t)p.F * p.L * p.K, p.K)) return 1;
    if (encode_rows_map(fn, &m->state, p.hE, (size_t)p.NB * p.L * p.K, p.K)) return 1;
    if (encode_rows_map(fn, &m->weights, p.model->dev_f16, (size_t)p.model->n_f16_blocks * 128, 128)) return 1;
    int dev = 0;
    CB2_CUDA(cudaGetDevice(&dev));
    CB2_CUDA(cudaDeviceGetAttribute(&p.num_sms, cudaDevAttrMultiProcessorCount, dev));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<4, EDGE_ENC_NODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(4, 2)));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<4, EDGE_DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(4, 2)));
    CB2_CUDA(cudaFuncSetAttribute(edge_tc_kernel<3, EDGE_ENC_EDGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(3, 3)));
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
    tp.mode = mode; tp.L = p.L; tp.K = p.K;
    tp.NPT = 128 / p.K < MAX_NPT ? 128 / p.K : MAX_NPT;
    if (tp.NPT < 1 || p.K % 8 != 0) tp.NPT = 1;      // node row blocks must start on a swizzle-atom (8-row) boundary
    tp.tiles_per_member = (p.L + tp.NPT - 1) / tp.NPT;
    tp.n_tiles = tp.tiles_per_member * p.NB;
    tp.lengths = p.lengths; tp.frame_of = p.frame_of; tp.nbr_idx = p.nbr_idx; tp.S = p.S;
    tp.mod_stride = mod_stride_b;
    const bool first = (layer == 0 && mode != EDGE_DEC);
    tp.in_is_frame = first ? 1 : 0;
    auto row_of = [&](const __half* w) { return (int)((w - m.dev_f16) / 128); };
    if (mode == EDGE_ENC_NODE) {
        const EncLayerW& e = m.enc[layer];
        tp.P = plan_P(p, 0); tp.Pc = p.Pc16[0]; tp.n_w = 2;
        tp.w_row[0] = row_of(e.W1b_h); tp.w_row[1] = row_of(e.W2_h); tp.b2 = e.b2;
    } else if (mode == EDGE_ENC_EDGE) {
        const EncLayerW& e = m.enc[layer];
        tp.P = plan_P(p, 1); tp.Pc = p.Pc16[1]; tp.n_w = 3;
        tp.w_row[0] = row_of(e.W11b_h); tp.w_row[1] = row_of(e.W12_h); tp.w_row[2] = row_of(e.W13_h);
        tp.b2 = e.b12; tp.b3 = e.b13;
        tp.mod = mod_base + CB2_MOD_ENC_OFF(layer);
        tp.res = reinterpret_cast<const __half*>(first ? p.hE0 : p.hE);
    } else {
        const DecLayerW& d = m.dec[layer];
        tp.P = plan_P(p, 0); tp.Pc = p.Pc16[0]; tp.n_w = 2;
        tp.w_row[0] = row_of(d.W1b2_h); tp.w_row[1] = row_of(d.W2_h); tp.b2 = d.b2;
    }
    if (mode == EDGE_ENC_EDGE) {
        const int grid = min(p.num_sms, (tp.n_tiles + 2) / 3);
        edge_tc_kernel<3, EDGE_ENC_EDGE><<<grid, 3 * 128, tc_smem_bytes(3, 3), s>>>(maps, tp);
    } else {
        const int grid = min(p.num_sms, (tp.n_tiles + 3) / 4);
        if (mode == EDGE_ENC_NODE) edge_tc_kernel<4, EDGE_ENC_NODE><<<grid, 4 * 128, tc_smem_bytes(4, 2), s>>>(maps, tp);
        else edge_tc_kernel<4, EDGE_DEC><<<grid, 4 * 128, tc_smem_bytes(4, 2), s>>>(maps, tp);
    }
    CB2_LAUNCH_CHECK();
    p.launches++;
    return 0;
}

}  // namespace cb2
