// tfce.hpp -- batched graph TFCE (tfce.cu): host-visible interface.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

namespace rmg {

// CSR adjacency (0-based) and vertex weights, all device-resident.
struct TfceGraph {
    const int *adj_ptr;      // V + 1
    const int *adj_idx;      // adj_ptr[V]
    const double *weights;   // V  (vertex area; ones by default); null -> every vertex weighs uniform_weight
    // connectivity != 0: the vertices are the voxels of an s0 x s1 x s2 grid (v = (i * s1 + j) * s2 + k) and the
    // neighbours are computed, not stored (6 / 18 / 26: Manhattan distance <= 1 / 2 / 3 inside the 3x3x3 box,
    // src/vertex_tfce.cpp:178-214); adj_ptr / adj_idx are then unused.
    int connectivity = 0;
    int s0 = 0, s1 = 0, s2 = 0;
    double uniform_weight = 1.0;
};

// The graph as the host entry points receive it: a CSR adjacency (+ optional weights), or a voxel grid (connectivity
// 6 / 18 / 26, dims in file order, every voxel weighing uniform_weight) whose neighbours are never materialised.
struct HostTfceGraph {
    const int32_t *adj_ptr = nullptr, *adj_idx = nullptr;
    const double *weights = nullptr;
    int connectivity = 0;
    int64_t dims[3] = {0, 0, 0};
    double uniform_weight = 1.0;
};

// Fills hg for the voxel grid `dims` (validated; RMG_ERR_INVALID + message otherwise).
int volume_graph(const int64_t *dims, int connectivity, double voxel_volume, HostTfceGraph &hg, char *err, size_t errlen);
// Argument checks of hg against V vertices.
int check_csr(const HostTfceGraph &hg, int64_t V, char *err, size_t errlen);
// Device image of hg (stream-ordered allocations returned through dptr / didx / dw for the caller to free).
int upload_graph(const HostTfceGraph &hg, int64_t V, cudaStream_t st, TfceGraph &g, int **dptr, int **didx, double **dw,
                 char *err, size_t errlen);

// Levels available when the thresholds come from a fixed step (volumes): levels are stored in one byte.
constexpr int kTfceStepLevels = 253;

// The voxel grid (file dimension order) as a CSR adjacency; connectivity 6, 18 or 26.  0, 1 (bad connectivity) or
// 2 (bad / too large dimensions).  Host only.
int grid_adjacency(const int64_t dims[3], int connectivity, std::vector<int32_t> &ptr, std::vector<int32_t> &idx);

// Bytes of scratch tfce_batch needs for nmaps maps of V vertices.
size_t tfce_workspace_bytes(int64_t V, int nmaps, int nsteps, int side);

// dmaps: nmaps maps of V doubles (map m at dmaps + m * ld); dout likewise (ldo).  side: 0 both, 1 positive, 2 negative.
// step > 0: thresholds are the multiples of `step` up to the map maximum (mincTFCE's d) and nsteps is ignored; returns
// cudaErrorNotSupported when a map needs more than kTfceStepLevels of them.
// Enqueued on st; synchronises the stream once (to read the number of threshold levels).
cudaError_t tfce_batch(const double *dmaps, int64_t V, int64_t ld, int nmaps, const TfceGraph &g, double E, double H,
                       int nsteps, int side, double *dout, int64_t ldo, void *workspace, cudaStream_t st, int *launches,
                       double step = 0.0);

}  // namespace rmg
