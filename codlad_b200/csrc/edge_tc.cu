// tcgen05 (bf16) tier of the per-edge message MLPs -- placeholder until the kernel lands.
#include "model.h"

namespace cb2 {

int edge_tc_prepare(Plan&) { set_error("bf16/tcgen05 tier is not built into this library yet"); return 1; }
void edge_tc_release(Plan&) {}
int launch_edge_tc(Plan&, int, int, const float*, int, cudaStream_t) { set_error("bf16/tcgen05 tier is not built"); return 1; }

}  // namespace cb2
