// Internal: packed VQ-VAE decode-side weights and launchers (decode.cu).
#pragma once
#include "model.h"

namespace cb2 {

struct VaeModel {
    int M = 4096;               // codebook size
    int angle_variant = 0;      // 0: IC_Decoder (N6), 1: IC_Decoder_angle (K3/K4)
    float cutoff = 21.0f;
    const float *codebook /* [M][3] */, *mean, *stdv, *mapw_t /* [3][36] */, *mapb;
    float* e2 = nullptr;        // [M] squared code norms
    const float* res_embed;     // [25][4]
    const float *inv0_t[4], *inv0_b[4], *inv1_t[4], *inv1_b[4];   // message_blocks.b.inv_dense.{0,1}
    const float *Wd_t /* [4][15][40] */, *bd /* [4][40] */;       // message_blocks.b.dist_embed.block.1
    const float *db1_t[4], *db1_b[4], *db3_t[4], *db3_b[4];       // dense_blocks.b.{1,3}
    const float *bb_dist, *sc_dist, *sa_embed;
    const float *ba1_t, *ba1_b, *ba3_t, *ba3_b, *bt1_t, *bt1_b, *bt3_t, *bt3_b;
    const float *sa1_t, *sa1_b, *sa3_t, *sa3_b;
    const float *tb1_t[4], *tb1_b[4], *tb3_t[4], *tb3_b[4];
    const float *ft1_t, *ft1_b, *ft3_t, *ft3_b;
    float* dev = nullptr;
};

int launch_codebook_norms(const float* cb, int M, float* e2, cudaStream_t s);
int launch_vq_lookup(const VaeModel& v, const float* x, int N, int L, const int* lengths, const int* frame_of, int denorm,
                     int* idx_out, float* zq_out, float* S40, cudaStream_t s);
int launch_ic_edge_filters(const VaeModel& v, const float* X, int F, int L, const int* row_ptr, const int* col, int E, float* w,
                           cudaStream_t s);
int launch_ic_decoder(const VaeModel& v, float* S40, float* phi, int N, int L, const int* frame_of, const int* lengths,
                      const int* cg_z, const int* row_ptr, const int* col, int E, const float* w, float* ic, cudaStream_t s,
                      long long* launches);
int launch_ic_to_xyz(const float* ca_full, const float* ic, int N, int L, const int* frame_of, const int* lengths,
                     const signed char* orders, const int* slot_atom, const long long* out_off, float* xyz, float* slots_dbg,
                     cudaStream_t s);

}  // namespace cb2
