// tfce.hpp -- batched graph TFCE (tfce.cu): host-visible interface.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace rmg {

// CSR adjacency (0-based) and vertex weights, all device-resident.
struct TfceGraph {
    const int *adj_ptr;      // V + 1
    const int *adj_idx;      // adj_ptr[V]
    const double *weights;   // V  (vertex area; ones by default)
};

// Bytes of scratch tfce_batch needs for nmaps maps of V vertices.
size_t tfce_workspace_bytes(int64_t V, int nmaps, int nsteps, int side);

// dmaps: nmaps maps of V doubles (map m at dmaps + m * ld); dout likewise (ldo).  side: 0 both, 1 positive, 2 negative.
// Enqueued on st; synchronises the stream once (to read the number of threshold levels).
cudaError_t tfce_batch(const double *dmaps, int64_t V, int64_t ld, int nmaps, const TfceGraph &g, double E, double H,
                       int nsteps, int side, double *dout, int64_t ldo, void *workspace, cudaStream_t st, int *launches);

}  // namespace rmg
