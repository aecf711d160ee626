// oracle/rstub/tfce_wrap.cpp -- TEST INFRASTRUCTURE.  C entry point around the reference's own graph_tfce /
// graph_tfce_wqu (src/vertex_tfce.cpp:92-157), compiled from /root/reference by `make ref` (oracle/Makefile).
#include <Rcpp.h>

#include <cstdint>
#include <cstdio>

std::vector<double> graph_tfce_wqu(std::vector<double> map, std::vector<std::vector<int> > adjacencies, double E, double H,
                                   int nsteps, std::vector<double> weights);
std::vector<double> graph_tfce(std::vector<double> map, std::vector<std::vector<int> > adjacencies, double E, double H,
                               int nsteps, std::vector<double> weights);

// side: 0 both, 1 positive, 2 negative  (vertexTFCE.numeric, R/minc_vertex_statistics.R:788-796)
extern "C" int ref_graph_tfce(const double *map, int64_t V, const int32_t *adj_ptr, const int32_t *adj_idx,
                              const double *weights, double E, double H, int nsteps, int side, double *out, char *err,
                              size_t errlen)
{
    try {
        std::vector<double> m(map, map + V), w(weights, weights + V);
        std::vector<std::vector<int> > adj(V);
        for (int64_t v = 0; v < V; ++v) adj[v].assign(adj_idx + adj_ptr[v], adj_idx + adj_ptr[v + 1]);
        std::vector<double> r;
        if (side == 1) {
            r = graph_tfce_wqu(m, adj, E, H, nsteps, w);
        } else if (side == 2) {
            for (auto &x : m) x = -x;
            r = graph_tfce_wqu(m, adj, E, H, nsteps, w);
            for (auto &x : r) x = -x;
        } else {
            r = graph_tfce(m, adj, E, H, nsteps, w);
        }
        for (int64_t v = 0; v < V; ++v) out[v] = r[v];
        return 0;
    } catch (const std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return 1;
    }
}

std::vector<std::vector<int> > neighbour_list(double x, double y, double z, int n);

// The reference's own neighbour_list (src/vertex_tfce.cpp:178-214), flattened to CSR.  Call with adj_idx == NULL to get
// the number of entries in *nnz (adj_ptr, x*y*z + 1 entries, is always filled).
extern "C" int ref_neighbour_list(int x, int y, int z, int n, int32_t *adj_ptr, int32_t *adj_idx, int64_t *nnz, char *err,
                                  size_t errlen)
{
    try {
        const std::vector<std::vector<int> > nl = neighbour_list(x, y, z, n);
        int64_t tot = 0;
        adj_ptr[0] = 0;
        for (size_t v = 0; v < nl.size(); ++v) {
            if (adj_idx)
                for (size_t k = 0; k < nl[v].size(); ++k) adj_idx[tot + (int64_t)k] = nl[v][k];
            tot += (int64_t)nl[v].size();
            adj_ptr[v + 1] = (int32_t)tot;
        }
        *nnz = tot;
        return 0;
    } catch (const std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return 1;
    }
}


std::vector<double> mesh_area(std::vector<double> vertices, std::vector<double> triangles);

// The reference's own mesh_area (src/vertex_tfce.cpp:221-257).
extern "C" int ref_mesh_area(const double *vertices, int64_t nv, const double *triangles, int64_t nt, double *areas)
{
    const std::vector<double> a = mesh_area(std::vector<double>(vertices, vertices + 3 * nv), std::vector<double>(triangles, triangles + 3 * nt));
    for (int64_t v = 0; v < nv; ++v) areas[v] = a[v];
    return 0;
}
