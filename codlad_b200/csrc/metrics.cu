// Sample-quality metrics of the evaluation step that follows the sampling path (reference
// utils/protein_module.py:245-364, test.py:168-188): bond-graph validity of a generated all-atom structure against its
// reference structure, and RMSD.  The reference builds two dense [Na, Na] distance matrices per structure on the CPU; here one
// kernel walks the pairs once, never materialises a matrix and returns six integer counts per structure.
//
// bond(i, j) = i != j and ||x_i - x_j|| < (r[z_i] + r[z_j]) * scale          (fp32, same operation order as the reference)
// counts per structure: {differing entries, reference bonds, generated bonds} over all atoms and over heavy atoms only
// (z != 1; dropping hydrogens from both structures = restricting the pairs to heavy-heavy ones).  Entries are counted over the
// full symmetric matrix, like `(bonds != ref_bonds).sum()` does.
#include "../../include/codlad_b200.h"
#include "common.cuh"

namespace cb2 {

namespace {

constexpr int MT_ROWS = 128;        // rows (atoms i) per CTA = threads per CTA
constexpr int MT_COLS = 256;        // atoms j staged per shared-memory tile

__global__ void __launch_bounds__(MT_ROWS) bond_graph_kernel(const float* __restrict__ xyz_ref, const float* __restrict__ xyz_gen,
                                                             const int* __restrict__ z, const long long* __restrict__ offsets,
                                                             const float* __restrict__ radius, int max_z, float scale,
                                                             unsigned long long* __restrict__ counts) {
    __shared__ float sR[MT_COLS][3], sG[MT_COLS][3], sRad[MT_COLS];
    __shared__ int sHeavy[MT_COLS];
    const int s = blockIdx.y;                                  // grid = (blocks of MT_ROWS atoms of the largest structure, structures)
    const long long base = offsets[s];
    const int na = (int)(offsets[s + 1] - base);
    if ((int)blockIdx.x * MT_ROWS >= na) return;               // smaller structure: nothing in this row block
    const int i = blockIdx.x * MT_ROWS + threadIdx.x;
    const bool live = i < na;
    float ri[3] = {0.f, 0.f, 0.f}, gi[3] = {0.f, 0.f, 0.f}, rad_i = 0.f;
    bool heavy_i = false;
    if (live) {
        for (int d = 0; d < 3; ++d) { ri[d] = xyz_ref[(base + i) * 3 + d]; gi[d] = xyz_gen[(base + i) * 3 + d]; }
        const int zi = z[base + i];
        rad_i = radius[min(max(zi, 0), max_z)];
        heavy_i = zi != 1;
    }
    unsigned int c[6] = {0u, 0u, 0u, 0u, 0u, 0u};        // diff, ref, gen (all atoms) | diff, ref, gen (heavy only)
    for (int j0 = 0; j0 < na; j0 += MT_COLS) {
        __syncthreads();
        for (int t = threadIdx.x; t < MT_COLS; t += MT_ROWS) {
            const int j = j0 + t;
            if (j < na) {
                for (int d = 0; d < 3; ++d) { sR[t][d] = xyz_ref[(base + j) * 3 + d]; sG[t][d] = xyz_gen[(base + j) * 3 + d]; }
                const int zj = z[base + j];
                sRad[t] = radius[min(max(zj, 0), max_z)];
                sHeavy[t] = zj != 1;
            }
        }
        __syncthreads();
        if (!live) continue;
        const int nj = min(MT_COLS, na - j0);
        for (int t = 0; t < nj; ++t) {
            if (j0 + t == i) continue;
            // (x_i - x_j).pow(2).sum(-1).sqrt() < (r_i + r_j) * scale, no fused multiply-add (the reference rounds every operation)
            const float cut = __fmul_rn(__fadd_rn(rad_i, sRad[t]), scale);
            float dr[3], dg[3];
            for (int d = 0; d < 3; ++d) { dr[d] = __fsub_rn(ri[d], sR[t][d]); dg[d] = __fsub_rn(gi[d], sG[t][d]); }
            const float d2r = __fadd_rn(__fadd_rn(__fmul_rn(dr[0], dr[0]), __fmul_rn(dr[1], dr[1])), __fmul_rn(dr[2], dr[2]));
            const float d2g = __fadd_rn(__fadd_rn(__fmul_rn(dg[0], dg[0]), __fmul_rn(dg[1], dg[1])), __fmul_rn(dg[2], dg[2]));
            const bool br = __fsqrt_rn(d2r) < cut, bg = __fsqrt_rn(d2g) < cut;
            const unsigned int hv = (heavy_i && sHeavy[t]) ? 1u : 0u;
            c[0] += br != bg; c[1] += br; c[2] += bg;
            c[3] += hv & (unsigned int)(br != bg); c[4] += hv & (unsigned int)br; c[5] += hv & (unsigned int)bg;
        }
    }
    // CTA reduction: warp shuffles, then one atomic per warp and counter (integer adds: order does not matter)
    for (int k = 0; k < 6; ++k) {
        unsigned int v = c[k];
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(counts + (size_t)s * 6 + k, (unsigned long long)v);
    }
}

// per structure: sum over atoms of ||x_gen - x_ref||^2 (all atoms, heavy atoms) in double, and the atom counts
__global__ void __launch_bounds__(256) rmsd_sums_kernel(const float* __restrict__ xyz_ref, const float* __restrict__ xyz_gen, const int* __restrict__ z,
                                                        const long long* __restrict__ offsets, double* __restrict__ out) {
    const int s = blockIdx.x;
    const long long base = offsets[s];
    const int na = (int)(offsets[s + 1] - base);
    double sa = 0.0, sh = 0.0, nh = 0.0;
    for (int i = threadIdx.x; i < na; i += blockDim.x) {
        double d2 = 0.0;
        for (int d = 0; d < 3; ++d) { const double t = (double)xyz_gen[(base + i) * 3 + d] - (double)xyz_ref[(base + i) * 3 + d]; d2 += t * t; }
        sa += d2;
        if (z[base + i] != 1) { sh += d2; nh += 1.0; }
    }
    __shared__ double red[3][8];
    double v[3] = {sa, sh, nh};
    for (int k = 0; k < 3; ++k) {
        for (int off = 16; off >= 1; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < 3; ++k) for (int w = 0; w < 8; ++w) t[k] += red[k][w];      // fixed order
        out[s * 4 + 0] = t[0]; out[s * 4 + 1] = (double)na; out[s * 4 + 2] = t[1]; out[s * 4 + 3] = t[2];
    }
}

// Minimum RMSD under rigid superposition (what md.rmsd computes for the diversity score, test.py:37-96), per structure pair: the
// 17 sums (centroids, inner products G_a / G_b, 3x3 cross-covariance) are accumulated in double over the atoms, then one thread takes
// the largest eigenvalue of Horn's 4x4 key matrix with cyclic Jacobi rotations: rmsd^2 = (G_a + G_b - 2 lambda_max) / N.
__global__ void __launch_bounds__(256) superposed_rmsd_kernel(const float* __restrict__ A, const float* __restrict__ B, const long long* __restrict__ offsets,
                                                              double* __restrict__ out) {
    const int s = blockIdx.x;
    const long long base = offsets[s];
    const int na = (int)(offsets[s + 1] - base);
    double acc[17];
    for (int k = 0; k < 17; ++k) acc[k] = 0.0;
    for (int i = threadIdx.x; i < na; i += blockDim.x) {
        double a[3], b[3];
        for (int d = 0; d < 3; ++d) { a[d] = (double)A[(base + i) * 3 + d]; b[d] = (double)B[(base + i) * 3 + d]; }
        for (int d = 0; d < 3; ++d) { acc[d] += a[d]; acc[3 + d] += b[d]; acc[6] += a[d] * a[d]; acc[7] += b[d] * b[d]; }
        for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) acc[8 + 3 * p + q] += a[p] * b[q];
    }
    __shared__ double red[17][8];
    for (int k = 0; k < 17; ++k) {
        double v = acc[k];
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double t[17];
    for (int k = 0; k < 17; ++k) { t[k] = 0.0; for (int w = 0; w < 8; ++w) t[k] += red[k][w]; }       // fixed order
    if (na == 0) { out[s] = 0.0; return; }
    const double n = (double)na;
    double S[3][3];
    for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) S[p][q] = t[8 + 3 * p + q] - t[p] * t[3 + q] / n;
    const double Ga = t[6] - (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) / n, Gb = t[7] - (t[3] * t[3] + t[4] * t[4] + t[5] * t[5]) / n;
    double K[4][4] = {
        {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
        {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
        {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
        {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += K[p][q] * K[p][q];
        if (off < 1e-30 * (Ga + Gb + 1e-300) * (Ga + Gb + 1e-300)) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (K[p][q] == 0.0) continue;
                const double theta = (K[q][q] - K[p][p]) / (2.0 * K[p][q]);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
                for (int k = 0; k < 4; ++k) { const double kp = K[k][p], kq = K[k][q]; K[k][p] = c * kp - sn * kq; K[k][q] = sn * kp + c * kq; }
                for (int k = 0; k < 4; ++k) { const double pk = K[p][k], qk = K[q][k]; K[p][k] = c * pk - sn * qk; K[q][k] = sn * pk + c * qk; }
            }
    }
    const double lam = fmax(fmax(K[0][0], K[1][1]), fmax(K[2][2], K[3][3]));
    out[s] = sqrt(fmax(0.0, (Ga + Gb - 2.0 * lam) / n));
}

}  // namespace

}  // namespace cb2

extern "C" int cb2_superposed_rmsd(const float* A, const float* B, const long long* offsets, int n_struct, double* out, void* stream) {
    using namespace cb2;
    if (!A || !B || !offsets || !out || n_struct < 0) { set_error("superposed_rmsd: bad argument"); return 1; }
    if (n_struct == 0) return 0;
    superposed_rmsd_kernel<<<n_struct, 256, 0, (cudaStream_t)stream>>>(A, B, offsets, out);
    CB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int cb2_eval_bond_graphs(const float* xyz_ref, const float* xyz_gen, const int* atomic_num, const long long* offsets, int n_struct,
                                    int max_atoms, const float* cov_radius, int max_z, float scale, long long* counts, double* sums, void* stream) {
    using namespace cb2;
    if (!xyz_ref || !xyz_gen || !atomic_num || !offsets || !cov_radius || !counts || !sums || n_struct < 0 || max_atoms < 0 || max_z < 1) {
        set_error("eval_bond_graphs: bad argument");
        return 1;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (n_struct > 65535) { set_error("eval_bond_graphs: at most 65535 structures per call"); return 1; }
    CB2_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_struct * 6 * sizeof(long long), s));
    if (n_struct == 0) return 0;
    if (max_atoms > 0) {
        bond_graph_kernel<<<dim3((max_atoms + MT_ROWS - 1) / MT_ROWS, n_struct), MT_ROWS, 0, s>>>(xyz_ref, xyz_gen, atomic_num, offsets, cov_radius, max_z, scale,
                                                                                                   reinterpret_cast<unsigned long long*>(counts));
        CB2_LAUNCH_CHECK();
    }
    rmsd_sums_kernel<<<n_struct, 256, 0, s>>>(xyz_ref, xyz_gen, atomic_num, offsets, sums);
    CB2_LAUNCH_CHECK();
    return 0;
}
