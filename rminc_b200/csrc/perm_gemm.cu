// perm_gemm.cu -- see perm_gemm.cuh for the design.
#include "perm_gemm.cuh"

#include <chrono>
#include <cstdio>

#include "async_ptx.cuh"
#include "capi_common.hpp"

namespace rmg {
namespace {

constexpr int NT = kPermNT, PITCH = kPermPitch;
// Subjects per pipeline stage and ring depth depend on the warp tile height: the A chunk of a stage is
// 4 * (KC/4) * WM * 32 doubles, so WM <= 6 affords 32-subject stages (3 of them), WM = 7..8 keeps 16 (4).

constexpr int WARPS_N = 2, WN = 4;  // warp tile: WM m8-tiles x 4 n8-tiles (32 voxels)

// Ring depth: as many stages as fit, at most 4.  (Tried in round 2: a 6-stage ring with the two warps of an SM
// sub-partition skewed by three stages, so that one could issue MMAs during the other's epilogue -- tensor pipe 85.9 %
// against 87.4 % in lockstep: the slower warp frees the stages, so the skew eats the producer's lead.)
constexpr int stages_for(int kp, int wm, int pg, int warps_m, int kc)
{
    const int stage_bytes = (warps_m * (kc / 4) * wm * 32 + kc * kPermPitch) * 8;
    const int epi_bytes = warps_m * pg * 8 * (kp + 1 + kPermMaxCols * kp) * 8;     // worst case: every column requested
    const int budget = (warps_m == 4 ? 226 : 112) * 1024 - 4 * 8 * 8 - warps_m * WARPS_N * 3 * 32 * 8 - epi_bytes;
    const int s = budget / stage_bytes;
    return s > 4 ? 4 : s < 2 ? 2 : s;
}

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 1/x for a positive, finite, normal x: MUFU.RCP64H seed (~20 bits) + two Newton steps = full
// double precision for about a quarter of the instructions of an IEEE division.
__device__ __forceinline__ double rcp_pos(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// ---- K5: one pass over Y for the per-voxel quantities every permutation shares.
__global__ void __launch_bounds__(256)
perm_prepass_kernel(const double *__restrict__ Y, int64_t V, int64_t ldY, int n, const double *__restrict__ mask,
                    double lo, double hi, int shift, double *__restrict__ y0, double *__restrict__ yy,
                    double *__restrict__ ys, int *__restrict__ any_zero_row)
{
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (v >= V) return;
    const bool two = v + 1 < V;
    bool a0 = true, a1 = two;
    if (mask) {
        a0 = mask_selects(mask[v], lo, hi);
        a1 = two && mask_selects(mask[v + 1], lo, hi);
    }
    double f0 = 0.0, f1 = 0.0, s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
    if (a0 || a1) {
        const double *yp = Y + v;
        if (shift) {
            f0 = yp[0];
            if (two) f1 = yp[1];
        }
        int j = 0;
        if (two) {  // V and ldY are even on this path: 16-byte loads
            for (; j + 4 <= n; j += 4) {
                double2 t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = ldg_stream2(yp + (int64_t)(j + u) * ldY);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double a = t[u].x - f0, b = t[u].y - f1;
                    s0 = fma(a, a, s0);
                    s1 = fma(b, b, s1);
                    t0 += a;
                    t1 += b;
                }
            }
        }
        for (; j < n; ++j) {
            const double a = yp[(int64_t)j * ldY] - f0;
            s0 = fma(a, a, s0);
            t0 += a;
            if (two) {
                const double b = yp[(int64_t)j * ldY + 1] - f1;
                s1 = fma(b, b, s1);
                t1 += b;
            }
        }
    }
    y0[v] = f0;
    yy[v] = a0 ? s0 : -1.0;
    ys[v] = t0;
    if (two) {
        y0[v + 1] = f1;
        yy[v + 1] = a1 ? s1 : -1.0;
        ys[v + 1] = t1;
    }
    if ((!a0 || (two && !a1)) && *any_zero_row == 0) *any_zero_row = 1;  // benign race: all writers store 1
}

// masked rows are 0 in every column and take part in the maximum (R/minc_voxel_statistics.R:1488)
__global__ void perm_floor_keys_kernel(unsigned long long *keys, int nk, const int *any_zero_row)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nk && *any_zero_row) atomicMax(&keys[i], ordered_key(0.0));
}

// The packed operand reaches the device through a kernel that reads the pinned host buffer directly (zero-copy over the
// host link) instead of a DMA copy: copies queued on the copy engines run in submission order, so a 3 MB operand
// batch submitted behind a 28 GB upload of Y would land only after it -- and the first voxel block could not start
// its permutations while the rest of Y is still on the link.
__global__ void __launch_bounds__(256)
perm_fetch_kernel(double2 *__restrict__ dst, const double2 *__restrict__ src_host, size_t n2)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src_host[i];
}

// An odd voxel count: the GEMM tiles cover an even number of voxels (16-byte bulk copies); the last voxel of the shard
// is handed to the explicit-residual kernel instead (unless it lies outside the mask).
__global__ void perm_flag_voxel_kernel(unsigned char *vflag, const double *yy, int64_t v)
{
    if (threadIdx.x == 0 && blockIdx.x == 0 && yy[v] >= 0.0) vflag[v] = 1;
}

// final: monotone proxy m = sign(b) b^2 / rss  ->  t = sign(m) sqrt(|m| * rdf / ct)
__global__ void perm_finalize_keys_kernel(unsigned long long *keys, const PermConst *pc, PermKernelParams P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.R * P.ncols) return;
    const int c = i / P.R, r = i % P.R;
    const double m = ordered_value(keys[i]);
    double t = m;
    if (isfinite(m)) {
        const double mag = sqrt(fabs(m) * pc[r].rdf / pc[r].ct[P.tcol[c]]);
        t = m < 0.0 ? -mag : mag;
        if (!isfinite(t)) t = 0.0;
    }
    keys[i] = ordered_key(t);
}

// Q_r[j][k] read back from the fragment-packed A operand (see the host packing in perm_gemm_run).
__device__ __forceinline__ double perm_q(const PermKernelParams &P, const PermConst &pc, int r, int k, int j)
{
    if (k < P.inv) return pc.c0;
    const int pt = r / P.perms_per_cta, rr = r % P.perms_per_cta;
    const int warp_m = rr / (P.pg * 8), pg = (rr / 8) % P.pg, g = rr % 8;
    const int mt = pg * P.kg + (k - P.inv);
    const int kc = j / 16, jj = j % 16, ks = jj / 4, q = jj % 4;
    return P.Apack[((size_t)pt * P.n_kc + kc) * P.a_chunk + (((size_t)warp_m * 4 + ks) * P.wm + mt) * 32 + g * 4 + q];
}

// ---- rare path: voxels for which some permutation's one-pass rss = |y'|^2 - |e'|^2 lost too
//      many digits (the permuted model explains > 99.9 % of the voxel).  The GEMM skips those
//      (permutation, voxel) pairs and raises vflag[v]; this kernel refits EVERY permutation of a
//      flagged voxel with explicit residuals (max is idempotent, so re-adding pairs the GEMM did
//      handle is harmless).  One CTA per flagged voxel, permutations across threads.
template <int KP>
__global__ void __launch_bounds__(128)
perm_exact_voxels_kernel(PermKernelParams P)
{
    __shared__ int hits[128];
    __shared__ int nhits;
    // scan 128 flags per block iteration; flagged voxels (normally none) are refitted one at a time
    for (int64_t base = (int64_t)blockIdx.x * 128; base < P.V; base += (int64_t)gridDim.x * 128) {
        if (threadIdx.x == 0) nhits = 0;
        __syncthreads();
        const int64_t mine = base + threadIdx.x;
        if (mine < P.V && P.vflag[mine]) hits[atomicAdd(&nhits, 1)] = threadIdx.x;
        __syncthreads();
        const int cnt = nhits;
      for (int hidx = 0; hidx < cnt; ++hidx) {
        const int64_t v = base + hits[hidx];
        const double *y = P.Y + v;
        const double y0 = P.y0[v];
        for (int r = threadIdx.x; r < P.R; r += blockDim.x) {
            const PermConst &pc = P.pc[r];
            double e[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) e[k] = 0.0;
            for (int j = 0; j < P.n; ++j) {
                const double yv = y[(int64_t)j * P.ldY] - y0;
#pragma unroll
                for (int k = 0; k < KP; ++k) e[k] = fma(perm_q(P, pc, r, k, j), yv, e[k]);
            }
            double rss = 0.0;
            for (int j = 0; j < P.n; ++j) {
                double fit = 0.0;
#pragma unroll
                for (int k = 0; k < KP; ++k) fit = fma(perm_q(P, pc, r, k, j), e[k], fit);
                const double res = (y[(int64_t)j * P.ldY] - y0) - fit;
                rss = fma(res, res, rss);
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) e[k] = fma(y0, pc.q1[k], e[k]);  // un-shift
            for (int c = 0; c < P.ncols; ++c) {
                const int ci = P.tcol[c];
                double b = 0.0;
#pragma unroll
                for (int j = 0; j < KP; ++j) b = fma(pc.Rinv[ci * 8 + j], e[j], b);
                double s = b * b / rss;
                if (!P.two_sided && b < 0.0) s = -s;
                if (!isfinite(s)) s = 0.0;
                atomicMax(&P.keys[(size_t)c * P.Rstride + r], ordered_key(s));
            }
        }
      }
        __syncthreads();
    }
}

// ---- the GEMM + fused statistics
// One pipeline stage (KC subjects = KC/4 k-steps) of MMAs for this warp.  The fragments of k-step
// ks+1 are loaded into a second register set before the MMAs of k-step ks are issued, so the
// LDS latency is hidden behind tensor-pipe work.  GUARD: the chunk may contain padding rows
// (j >= n) whose shared-memory rows hold stale data and must read as 0.
template <int WM, int KC, bool GUARD, bool PREFETCH>
__device__ __forceinline__ void consume_chunk(double (&acc)[WM][WN][2], const double *__restrict__ As,
                                              const double *__restrict__ Bs, int ksteps, int j0, int q, int n)
{
    constexpr int KS = KC / 4;
    double a[2][WM], b[2][WN];
    auto load = [&](int buf, int ks) {
#pragma unroll
        for (int mt = 0; mt < WM; ++mt) a[buf][mt] = As[(ks * WM + mt) * 32];
        const bool real = !GUARD || (j0 + ks * 4 + q) < n;
#pragma unroll
        for (int nt = 0; nt < WN; ++nt) {
            const double t = Bs[ks * 4 * PITCH + nt * 8];
            b[buf][nt] = real ? t : 0.0;
        }
    };
    if (PREFETCH) load(0, 0);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        if (!GUARD || ks < ksteps) {
            if (!PREFETCH) load(ks & 1, ks);
            else if (ks + 1 < KS && (!GUARD || ks + 1 < ksteps)) load((ks + 1) & 1, ks + 1);
#pragma unroll
            for (int mt = 0; mt < WM; ++mt)
#pragma unroll
                for (int nt = 0; nt < WN; ++nt)
                    dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[ks & 1][mt], b[ks & 1][nt]);
        }
    }
}

// KP design columns; INV = 1 when component 0 of every permuted design is the (permutation-
// invariant) constant column c0 * 1: its e_0 = c0 * sum_j y_j comes from the prepass and only
// KG = KP - INV components per permutation go through the tensor cores.
// TWO: alternative = "two.sided" (the default): every statistic is >= 0, so pairs without a contribution (outside the
// mask, past V, handed to the exact path, padding permutations) may simply contribute 0 and the statistics loop is
// straight-line code.  The one-sided variant keeps them apart (a 0 must not beat an all-negative column).
template <int KP, int PG, int INV, int KC, int STAGES, bool TWO, int WARPS_M>
__global__ void __launch_bounds__((WARPS_M * WARPS_N + 1) * 32, WARPS_M == 4 ? 1 : 2)
perm_gemm_kernel(PermKernelParams P)
{
    constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
    constexpr int KG = KP - INV;
    constexpr int WM = KG * PG;                           // m8-tiles per warp
    constexpr int A_CHUNK = WARPS_M * (KC / 4) * WM * 32; // doubles per stage
    constexpr int B_CHUNK = KC * PITCH;
    constexpr int PERMS_PER_CTA = WARPS_M * PG * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);
    double *sB = sA + (size_t)STAGES * A_CHUNK;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + (size_t)STAGES * B_CHUNK);
    uint64_t *empty = full + STAGES;
    uint64_t *epi_full = empty + STAGES, *epi_empty = epi_full + 1;
    // per-voxel constants (|y'|^2, y0, sum y') of each warp's 32 voxels of the current voxel tile, staged by the warp
    // itself (no CTA-wide barrier: the warps of a CTA deliberately do not run in lockstep)
    double *sv_all = reinterpret_cast<double *>(epi_empty + 1);    // [CONSUMER_WARPS][3][32]
    // the epilogue's per-permutation constants of the current permutation tile (q1, c0, the requested rows of R^-1),
    // bulk-copied by the producer while the tile's MMAs run: no global load is left in the epilogue
    double *sE = sv_all + CONSUMER_WARPS * 96;                     // [PERMS_PER_CTA][epi_stride]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        mbar_init(epi_full, 1);
        mbar_init(epi_empty, CONSUMER_WARPS);
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t n_vt = (P.V + NT - 1) / NT;
    const int n = P.n;

    if (warp == CONSUMER_WARPS) {
        // ===== producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, epi_phase = 0;
            const uint32_t epi_bytes = (uint32_t)(PERMS_PER_CTA * P.epi_stride * 8);
            // The constants of tile pt replace those of tile pt - 1 once every consumer warp has left that tile's
            // epilogue (epi_empty).  Issued a few stages into the tile, so the wait never holds up the ring refill
            // that runs during the epilogue.
            const int kc_epi = min(STAGES, P.n_kc - 1);
            for (int64_t vt = blockIdx.x; vt < n_vt; vt += gridDim.x) {
                const int64_t v0 = vt * NT;
                const int cols_here = (int)min((int64_t)NT, P.V - v0);
                const uint32_t row_bytes = (uint32_t)cols_here * 8u;
                for (int pt = 0; pt < P.n_pt; ++pt) {
                    for (int kc = 0; kc < P.n_kc; ++kc) {
                        if (kc == kc_epi) {
                            mbar_wait(epi_empty, epi_phase ^ 1);
                            mbar_expect_tx(epi_full, epi_bytes);
                            bulk_load_1d(sE, P.epi + (size_t)pt * PERMS_PER_CTA * P.epi_stride, epi_bytes, epi_full);
                            epi_phase ^= 1;
                        }
                        const int j0 = kc * KC;
                        const int rows = min(KC, n - j0);
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], (uint32_t)(A_CHUNK * 8) + rows * row_bytes);
                        bulk_load_1d(sA + (size_t)stage * A_CHUNK,
                                     P.Apack + ((size_t)pt * P.n_kc + kc) * A_CHUNK, A_CHUNK * 8, &full[stage]);
                        double *bdst = sB + (size_t)stage * B_CHUNK;
                        const double *bsrc = P.Y + v0 + (int64_t)j0 * P.ldY;
                        for (int r = 0; r < rows; ++r)
                            bulk_load_1d(bdst + r * PITCH, bsrc + (int64_t)r * P.ldY, row_bytes, &full[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        return;
    }

    // ===== consumers
    const int warp_m = warp / WARPS_N, warp_n = warp % WARPS_N;
    const int g = lane >> 2, q = lane & 3;
    uint32_t stage = 0, phase = 0, epi_phase = 0;
    const bool tail_guard = (n % KC) != 0;
    const int estride = P.epi_stride;

    double *sv = sv_all + warp * 96;                          // [3][32]
    // this lane's voxel of the warp's 32: its per-voxel sums are fetched one voxel tile ahead
    double pf_yy = -1.0, pf_y0 = 0.0, pf_ys = 0.0;
    auto prefetch = [&](int64_t vt) {
        const int64_t v = vt * NT + warp_n * 32 + lane;
        const bool ok = vt < n_vt && v < P.V;
        pf_yy = ok ? P.yy[v] : -1.0;                   // -1: no contribution (outside the mask or past V)
        pf_y0 = ok ? P.y0[v] : 0.0;
        pf_ys = (ok && INV) ? P.ys[v] : 0.0;
    };
    prefetch(blockIdx.x);

    for (int64_t vt = blockIdx.x; vt < n_vt; vt += gridDim.x) {
        const int64_t vbase = vt * NT + warp_n * 32 + 2 * q;  // + nt*8 + h
        const int cbase = 2 * q;                              // column of this thread inside the warp's 32 voxels
        __syncwarp();                                         // the previous tile's epilogue has read sv
        sv[lane] = pf_yy;
        sv[32 + lane] = pf_y0;
        sv[64 + lane] = pf_ys;
        __syncwarp();
        prefetch(vt + gridDim.x);
        for (int pt = 0; pt < P.n_pt; ++pt) {
            double acc[WM][WN][2];
#pragma unroll
            for (int mt = 0; mt < WM; ++mt)
#pragma unroll
                for (int nt = 0; nt < WN; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

            for (int kc = 0; kc < P.n_kc; ++kc) {
                const int j0 = kc * KC;
                mbar_wait(&full[stage], phase);
                const double *As = sA + (size_t)stage * A_CHUNK + (size_t)warp_m * (KC / 4) * WM * 32 + lane;
                const double *Bs = sB + (size_t)stage * B_CHUNK + q * PITCH + warp_n * 32 + g;
                if (tail_guard && kc == P.n_kc - 1)
                    consume_chunk<WM, KC, true, false>(acc, As, Bs, (n - j0 + 3) / 4, j0, q, n);
                else
                    consume_chunk<WM, KC, false, false>(acc, As, Bs, KC / 4, j0, q, n);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }

            mbar_wait(epi_full, epi_phase);
            epi_phase ^= 1;
            // ---- fused epilogue: this thread holds, for PG permutations, the KG GEMM components of
            //      e = Q_r'y for 8 voxels (nt = 0..3, h = 0..1).
#pragma unroll
            for (int pg = 0; pg < PG; ++pg) {
                const int r = pt * PERMS_PER_CTA + (warp_m * PG + pg) * 8 + g;
                const bool rvalid = r < P.R;
                const double *ep = sE + (size_t)((warp_m * PG + pg) * 8 + g) * estride;   // zeros for padding permutations
                // w > 0: 1/rss of a regular fit; 0: the statistic counts as 0 (non-finite in the reference; for TWO also
                // every pair without a contribution); < 0 (one-sided only): no contribution
                double w[WN][2];
                double e0[WN][2];   // INV: the un-shifted invariant component
                {
                    double q1[KP];
#pragma unroll
                    for (int k = 0; k < KP; ++k) q1[k] = ep[k];
                    const double c0 = INV ? ep[KP] : 0.0;
#pragma unroll
                    for (int nt = 0; nt < WN; ++nt) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int col = cbase + nt * 8 + h;
                            const double yy = sv[col];
                            const double y0 = sv[32 + col];
                            double ee = 0.0;
                            e0[nt][h] = 0.0;
                            if (INV) {
                                const double es = c0 * sv[64 + col];   // c0 * sum(y - y0)
                                ee = es * es;
                                e0[nt][h] = fma(y0, q1[0], es);
                            }
#pragma unroll
                            for (int k = 0; k < KG; ++k) {
                                const double es = fma(-y0, q1[k + INV], acc[pg * KG + k][nt][h]);
                                ee = fma(es, es, ee);
                            }
                            const double rss = yy - ee;
                            const bool ok = rvalid && yy >= 0.0;                       // NaN sums fail this too
                            const bool exact = ok && yy > 0.0 && rss <= kExactThreshold * yy;
                            if (exact) P.vflag[vbase + nt * 8 + h] = 1;  // model explains ~everything: explicit residuals later
                            const bool good = ok && !exact && rss > 1e-290 && rss < 1e290;
                            const double inv = rcp_pos(good ? rss : 1.0);
                            w[nt][h] = good ? inv : (TWO || (ok && !exact)) ? 0.0 : -1.0;
                        }
                    }
                }
#pragma unroll 1
                for (int c = 0; c < P.ncols; ++c) {
                    double rrow[KP];
#pragma unroll
                    for (int j = 0; j < KP; ++j) rrow[j] = ep[KP + 1 + c * KP + j];
                    double best = TWO ? 0.0 : -INFINITY;
#pragma unroll
                    for (int nt = 0; nt < WN; ++nt) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            double b = INV ? rrow[0] * e0[nt][h] : 0.0;
#pragma unroll
                            for (int j = 0; j < KG; ++j) b = fma(rrow[j + INV], acc[pg * KG + j][nt][h], b);
                            const double ww = w[nt][h];
                            double s = b * b * fabs(ww);
                            if (!TWO) s = b < 0.0 ? -s : s;
                            s = fabs(s) < INFINITY ? s : 0.0;          // non-finite -> 0 (R/minc_voxel_statistics.R:1478)
                            if (TWO)
                                best = s > best ? s : best;
                            else
                                best = (ww >= 0.0 && s > best) ? s : best;
                        }
                    }
                    // the 4 lanes of a quad hold different voxels of the same permutation
                    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, 1));
                    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, 2));
                    if (q == 0 && rvalid && best > -INFINITY) atomicMax(&P.keys[(size_t)c * P.Rstride + r], ordered_key(best));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(epi_empty);     // this warp is done with the tile's constants
        }
    }
}

template <int KP, int PG, int INV, int KC, int STAGES, bool TWO, int WARPS_M>
cudaError_t launch_gemm_cfg(const PermKernelParams &P, int sm_count, cudaStream_t st)
{
    constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
    constexpr int WM = (KP - INV) * PG;
    constexpr int A_CHUNK = WARPS_M * (KC / 4) * WM * 32;
    constexpr int PERMS_PER_CTA = WARPS_M * PG * 8;
    const size_t smem = (size_t)STAGES * (A_CHUNK + KC * PITCH) * 8 + (2 * STAGES + 2) * 8 + CONSUMER_WARPS * 96 * 8 +
                        (size_t)PERMS_PER_CTA * P.epi_stride * 8;
    static_assert((size_t)STAGES * (A_CHUNK + KC * PITCH) * 8 + (2 * STAGES + 2) * 8 + CONSUMER_WARPS * 96 * 8 +
                      (size_t)PERMS_PER_CTA * (KP + 1 + kPermMaxCols * KP) * 8 <= 227 * 1024, "ring does not fit");
    auto kern = perm_gemm_kernel<KP, PG, INV, KC, STAGES, TWO, WARPS_M>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t n_vt = (P.V + NT - 1) / NT;
    const int grid = (int)std::min<int64_t>(n_vt, (int64_t)sm_count * (WARPS_M == 4 ? 1 : 2));
    kern<<<grid, (CONSUMER_WARPS + 1) * 32, smem, st>>>(P);
    return cudaGetLastError();
}

template <int KP, int PG, int INV>
cudaError_t launch_gemm(const PermKernelParams &P, int sm_count, cudaStream_t st, int warps_m)
{
    constexpr int WM = (KP - INV) * PG;
    constexpr int S4 = stages_for(KP, WM, PG, 4, 16), S2 = stages_for(KP, WM, PG, 2, 16);
    if (P.two_sided)
        return warps_m == 4 ? launch_gemm_cfg<KP, PG, INV, 16, S4, true, 4>(P, sm_count, st)
                            : launch_gemm_cfg<KP, PG, INV, 16, S2, true, 2>(P, sm_count, st);
    return warps_m == 4 ? launch_gemm_cfg<KP, PG, INV, 16, S4, false, 4>(P, sm_count, st)
                        : launch_gemm_cfg<KP, PG, INV, 16, S2, false, 2>(P, sm_count, st);
}

// permutation groups (of 8) per warp for KG GEMM components per permutation: KG * PG <= 8 m-tiles
constexpr int pg_for(int kg) { return kg == 1 ? 8 : kg == 2 ? 4 : kg <= 4 ? 2 : 1; }

}  // namespace

bool perm_gemm_supported(int n, int p, int64_t V, int64_t ldY, const double *dY, const int32_t *hcols, int ncols)
{
    if (ncols < 1 || ncols > kPermMaxCols) return false;
    for (int c = 0; c < ncols; ++c)
        if (hcols[c] < p + 2 || hcols[c] > 2 * p + 1) return false;  // fused epilogue covers the tvalue- columns
    // an odd V is fine (the last voxel takes the explicit-residual kernel); rows must stay 16-byte aligned
    return p >= 1 && p <= 8 && n >= 4 && V >= 2 && ldY % 2 == 0 && V < 0x7fffffffLL &&
           (reinterpret_cast<uintptr_t>(dY) & 15) == 0;
}

// ---- host side: geometry + packing
cudaError_t perm_pack_plan(const PermSet &ps, int n, int p, const int32_t *hcols, int ncols, PermPack &pk)
{
    const int R = ps.R, KP = p;
    pk.R = R; pk.n = n; pk.p = p;
    pk.ncols = ncols;
    for (int c = 0; c < ncols && c < kPermMaxCols; ++c) pk.tcol[c] = hcols[c] - (p + 2);
    pk.epi_stride = p + 1 + ncols * p;
    pk.shift = ps.all_intercept ? 1 : 0;
    // Is component 0 of every permuted design the same constant column (the intercept, first in
    // pivot order)?  Then e_0 = c0 * sum(y) is permutation-invariant and needs no tensor-core row.
    int INV = (KP >= 2 && ps.all_intercept && !std::getenv("RMG_PERM_NO_INV")) ? 1 : 0;
    pk.c0.assign(R, 0.0);
    auto constant_column = [&](const Design &D, double &value) {
        if (D.k < 1) return false;
        const double *q = D.Q.data();
        double lo = q[0], hi = q[0], sum = 0.0;
        for (int j = 0; j < n; ++j) { lo = std::min(lo, q[j]); hi = std::max(hi, q[j]); sum += q[j]; }
        value = sum / n;
        return (hi - lo <= 1e-13 * std::fabs(hi)) && hi != 0.0;
    };
    double base_c0 = 0.0;
    const bool base_const = constant_column(ps.base, base_c0);   // a row gather keeps a constant column constant
    for (int r = 0; r < R && INV; ++r) {
        if (ps.is_perm(r)) {
            if (!base_const) INV = 0;
            pk.c0[r] = base_c0;
        } else if (!constant_column(ps.own[r], pk.c0[r])) {
            INV = 0;
        }
    }
    pk.INV = INV;
    pk.KG = KP - INV;
    pk.PG = pg_for(pk.KG);
    pk.WM = pk.KG * pk.PG;
    pk.KC = 16;
    // 4 x 2 consumer warps, one CTA per SM.  (2 x 2 warps with two CTAs per SM measured the same
    // 1.30e3 perms/s but re-reads Y more: 35 GB vs 28.5 GB of DRAM traffic per 64 permutations.)
    pk.WARPS_M = 4;
    if (const char *e = std::getenv("RMG_PERM_CTAS")) pk.WARPS_M = std::atoi(e) == 2 ? 2 : 4;
    pk.A_CHUNK = pk.WARPS_M * (pk.KC / 4) * pk.WM * 32;
    pk.perms_per_cta = pk.WARPS_M * pk.PG * 8;
    pk.n_pt = (R + pk.perms_per_cta - 1) / pk.perms_per_cta;
    pk.n_kc = (n + pk.KC - 1) / pk.KC;
    pk.tiles_per_batch = std::max(1, 256 / pk.perms_per_cta);
    pk.first_tiles = 1;
    pk.nbatch = pk.n_pt <= pk.first_tiles ? 1 : 1 + (pk.n_pt - pk.first_tiles + pk.tiles_per_batch - 1) / pk.tiles_per_batch;
    pk.apack_doubles = (size_t)pk.n_pt * pk.n_kc * pk.A_CHUNK;
    // pinned staging (leased from the process-wide pool: the DMA engines of every device read it directly)
    const size_t bytesA = (pk.apack_doubles * 8 + 255) & ~size_t(255);
    const size_t bytesP = ((size_t)std::max(R, 1) * sizeof(PermConst) + 255) & ~size_t(255);
    const size_t bytesE = (size_t)pk.n_pt * pk.perms_per_cta * pk.epi_stride * 8;
    char *base = static_cast<char *>(pinned_acquire(bytesA + bytesP + bytesE));
    if (!base) return cudaErrorMemoryAllocation;
    pk.lease = base;
    pk.hA = reinterpret_cast<double *>(base);
    pk.hpc = reinterpret_cast<PermConst *>(base + bytesA);
    pk.hE = reinterpret_cast<double *>(base + bytesA + bytesP);
    return cudaSuccess;
}

void perm_pack_release(PermPack &pk)
{
    if (pk.lease) pinned_release(pk.lease);
    pk.lease = nullptr;
    pk.hA = nullptr;
    pk.hpc = nullptr;
    pk.hE = nullptr;
}

void perm_pack_batch(const PermSet &ps, PermPack &pk, int b)
{
    const int n = pk.n, p = pk.p, INV = pk.INV, KG = pk.KG, PG = pk.PG, WM = pk.WM, KC = pk.KC, A_CHUNK = pk.A_CHUNK;
    const int perms_per_cta = pk.perms_per_cta, n_kc = pk.n_kc;
    const int pt0 = pk.batch_pt0(b), pt1 = pk.batch_pt1(b);
    const int r0 = pk.batch_r0(b), r1 = pk.batch_r1(b);
    double *Apack = pk.hA;
    PermConst *pc = pk.hpc;
    std::memset(Apack + (size_t)pt0 * n_kc * A_CHUNK, 0, (size_t)(pt1 - pt0) * n_kc * A_CHUNK * sizeof(double));   // padding
    const int es = pk.epi_stride;
    std::memset(pk.hE + (size_t)pt0 * perms_per_cta * es, 0, (size_t)(pt1 - pt0) * perms_per_cta * es * sizeof(double));
    parallel_for(r1 - r0, [&](int i0, int i1) {
        for (int r = r0 + i0; r < r0 + i1; ++r) {
            const Design &D = ps.design(r);
            const bool gather = ps.is_perm(r);
            const int32_t *ix = ps.idx + (size_t)r * n;
            PermConst &c = pc[r];
            std::memset(&c, 0, sizeof c);
            for (int i = 0; i < p; ++i) {
                c.q1[i] = D.q1[i];
                c.ct[i] = D.ct[i];
                for (int j = 0; j < p; ++j) c.Rinv[i * 8 + j] = D.Rinv[i + (size_t)j * p];
            }
            c.rdf = (double)(n - p);
            c.c0 = pk.c0[r];
            double *e = pk.hE + (size_t)r * es;
            for (int i = 0; i < p; ++i) e[i] = D.q1[i];
            e[p] = pk.c0[r];
            for (int cc = 0; cc < pk.ncols; ++cc)
                for (int j = 0; j < p; ++j) e[p + 1 + cc * p + j] = D.Rinv[pk.tcol[cc] + (size_t)j * p];
            const int pt = r / perms_per_cta, rr = r % perms_per_cta;
            const int warp_m = rr / (PG * 8), pg = (rr / 8) % PG, g = rr % 8;
            for (int k = INV; k < D.k; ++k) {
                const double *qcol = D.Q.data() + (size_t)k * n;
                const int mt = pg * KG + (k - INV);
                for (int j = 0; j < n; ++j) {
                    const int kc = j / KC, jj = j % KC, ks = jj / 4, q = jj % 4;
                    const int lane = g * 4 + q;
                    // [pt][kc] chunk, inside: [warp_m][ks][mt][lane]  (one MMA A-fragment = 32 consecutive doubles)
                    Apack[(((size_t)pt * n_kc + kc) * A_CHUNK) + (((size_t)warp_m * (KC / 4) + ks) * WM + mt) * 32 + lane] =
                        gather ? qcol[ix[j]] : qcol[j];
                }
            }
        }
    });
}

// ---- device side
cudaError_t perm_dev_begin(PermDevice &d, const PermPack &pk, int device, int sm_count, const PermGemmArgs &a, cudaStream_t st,
                           const PermPre *resident_pre, int *launches)
{
    d.device = device;
    d.sm_count = sm_count;
    d.st = st;
    d.V = a.V;
    d.own_pre = resident_pre == nullptr;
    const int R = pk.R;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    const size_t oA = carve(pk.apack_doubles * 8), oP = carve((size_t)R * sizeof(PermConst)), oVf = carve((size_t)a.V);
    const size_t oE = carve((size_t)pk.n_pt * pk.perms_per_cta * pk.epi_stride * 8);
    size_t oY0 = 0, oYY = 0, oYS = 0, oFlag = 0;
    if (d.own_pre) {
        oY0 = carve((size_t)a.V * 8); oYY = carve((size_t)a.V * 8); oYS = carve((size_t)a.V * 8); oFlag = carve(256);
    }
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&d.ws), off, st);
    if (e != cudaSuccess) return e;
    if (d.own_pre) {
        d.pre.y0 = reinterpret_cast<double *>(d.ws + oY0);
        d.pre.yy = reinterpret_cast<double *>(d.ws + oYY);
        d.pre.ys = reinterpret_cast<double *>(d.ws + oYS);
        d.pre.any_zero = reinterpret_cast<int *>(d.ws + oFlag);
        if ((e = cudaMemsetAsync(d.ws + oFlag, 0, 256, st)) != cudaSuccess) return e;
    } else {
        d.pre = *resident_pre;
    }
    if ((e = cudaMemsetAsync(d.ws + oVf, 0, (size_t)a.V, st)) != cudaSuccess) return e;   // exact-path flags of THIS call

    PermKernelParams &P = d.P;
    std::memset(&P, 0, sizeof P);
    P.Y = a.Y; P.V = a.V; P.ldY = a.ldY; P.n = pk.n; P.p = pk.p; P.R = R; P.Rstride = R; P.ncols = a.ncols;
    P.two_sided = a.two_sided;
    P.shift = pk.shift;
    P.Apack = reinterpret_cast<double *>(d.ws + oA);
    P.pc = reinterpret_cast<PermConst *>(d.ws + oP);
    P.y0 = d.pre.y0; P.yy = d.pre.yy; P.ys = d.pre.ys;
    P.keys = a.keys;
    P.inv = pk.INV; P.kg = pk.KG; P.pg = pk.PG; P.wm = pk.WM; P.warps_m = pk.WARPS_M; P.a_chunk = pk.A_CHUNK;
    P.perms_per_cta = pk.perms_per_cta;
    P.vflag = reinterpret_cast<unsigned char *>(d.ws + oVf);
    P.n_pt = pk.n_pt;
    P.n_kc = pk.n_kc;
    P.epi = reinterpret_cast<double *>(d.ws + oE);
    P.epi_stride = pk.epi_stride;
    for (int c = 0; c < a.ncols; ++c) P.tcol[c] = pk.tcol[c];
    d.ev.assign(pk.nbatch, nullptr);
    if ((e = cudaStreamCreateWithFlags(&d.cs, cudaStreamNonBlocking)) != cudaSuccess) return e;
    (void)launches;
    return cudaSuccess;
}

cudaError_t perm_dev_prepass(PermDevice &d, const PermGemmArgs &a, int64_t v0, int64_t cv, int *launches)
{
    if (cv <= 0) return cudaSuccess;
    const int64_t pairs = (cv + 1) / 2;
    perm_prepass_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, d.st>>>(a.Y + v0, cv, a.ldY, a.n, a.mask ? a.mask + v0 : nullptr,
                                                                            a.mask_lo, a.mask_hi, d.P.shift, d.pre.y0 + v0,
                                                                            d.pre.yy + v0, d.pre.ys + v0, d.pre.any_zero);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t perm_dev_upload(PermDevice &d, const PermPack &pk, int b)
{
    const int pt0 = pk.batch_pt0(b), pt1 = pk.batch_pt1(b);
    const int r0 = pk.batch_r0(b), r1 = pk.batch_r1(b);
    const size_t offA = (size_t)pt0 * pk.n_kc * pk.A_CHUNK, nA = (size_t)(pt1 - pt0) * pk.n_kc * pk.A_CHUNK;
    static_assert(sizeof(PermConst) % 16 == 0, "PermConst is fetched in 16-byte pieces");
    void *devA = nullptr, *devP = nullptr;       // device-side addresses of the pinned host staging
    cudaError_t e = cudaHostGetDevicePointer(&devA, pk.hA, 0);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&devP, pk.hpc, 0);
    if (e != cudaSuccess) return e;
    void *devE = nullptr;
    if ((e = cudaHostGetDevicePointer(&devE, pk.hE, 0)) != cudaSuccess) return e;
    const size_t offE = (size_t)pt0 * pk.perms_per_cta * pk.epi_stride, nE2 = (size_t)(pt1 - pt0) * pk.perms_per_cta * pk.epi_stride / 2;
    if (nE2)
        perm_fetch_kernel<<<(unsigned)std::min<size_t>(16, (nE2 + 255) / 256), 256, 0, d.cs>>>(
            reinterpret_cast<double2 *>(const_cast<double *>(d.P.epi) + offE),
            reinterpret_cast<const double2 *>(static_cast<double *>(devE) + offE), nE2);
    const size_t nA2 = nA / 2, nP2 = (size_t)(r1 - r0) * sizeof(PermConst) / 16;
    if (nA2) {
        const unsigned blocks = (unsigned)std::min<size_t>(64, (nA2 + 255) / 256);
        perm_fetch_kernel<<<blocks, 256, 0, d.cs>>>(reinterpret_cast<double2 *>(const_cast<double *>(d.P.Apack) + offA),
                                                    reinterpret_cast<const double2 *>(static_cast<double *>(devA) + offA), nA2);
    }
    if (nP2) {
        const unsigned blocks = (unsigned)std::min<size_t>(16, (nP2 + 255) / 256);
        perm_fetch_kernel<<<blocks, 256, 0, d.cs>>>(reinterpret_cast<double2 *>(const_cast<PermConst *>(d.P.pc) + r0),
                                                    reinterpret_cast<const double2 *>(static_cast<PermConst *>(devP) + r0), nP2);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess && !d.ev[b]) e = cudaEventCreateWithFlags(&d.ev[b], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(d.ev[b], d.cs);
    return e;
}

cudaError_t perm_dev_launch(PermDevice &d, const PermPack &pk, int b0, int b1, int64_t v0, int64_t cv, int *launches)
{
    if (cv <= 0 || b1 <= b0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (cv & 1) {
        perm_flag_voxel_kernel<<<1, 32, 0, d.st>>>(d.P.vflag, d.P.yy, v0 + cv - 1);
        *launches += 1;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (--cv == 0) return cudaSuccess;
    }
    for (int b = b0; b < b1; ++b)
        if ((e = cudaStreamWaitEvent(d.st, d.ev[b], 0)) != cudaSuccess) return e;
    const int pt0 = pk.batch_pt0(b0), pt1 = pk.batch_pt1(b1 - 1);
    const int r0 = pk.batch_r0(b0), r1 = pk.batch_r1(b1 - 1);
    PermKernelParams Pb = d.P;
    Pb.Y = d.P.Y + v0;
    Pb.V = cv;
    Pb.y0 = d.P.y0 + v0;
    Pb.yy = d.P.yy + v0;
    Pb.ys = d.P.ys + v0;
    Pb.vflag = d.P.vflag + v0;
    Pb.Apack = d.P.Apack + (size_t)pt0 * pk.n_kc * pk.A_CHUNK;
    Pb.pc = d.P.pc + r0;
    Pb.epi = d.P.epi + (size_t)pt0 * pk.perms_per_cta * pk.epi_stride;
    Pb.keys = d.P.keys + r0;
    Pb.R = r1 - r0;
    Pb.n_pt = pt1 - pt0;
    switch (pk.p * 2 + pk.INV) {
#define RMG_PG_CASE(K) case K * 2: e = launch_gemm<K, pg_for(K), 0>(Pb, d.sm_count, d.st, pk.WARPS_M); break;
#define RMG_PG_CASE1(K) case K * 2 + 1: e = launch_gemm<K, pg_for(K - 1), 1>(Pb, d.sm_count, d.st, pk.WARPS_M); break;
        RMG_PG_CASE(1) RMG_PG_CASE(2) RMG_PG_CASE(3) RMG_PG_CASE(4) RMG_PG_CASE(5) RMG_PG_CASE(6) RMG_PG_CASE(7) RMG_PG_CASE(8)
        RMG_PG_CASE1(2) RMG_PG_CASE1(3) RMG_PG_CASE1(4) RMG_PG_CASE1(5) RMG_PG_CASE1(6) RMG_PG_CASE1(7) RMG_PG_CASE1(8)
#undef RMG_PG_CASE
#undef RMG_PG_CASE1
    default: e = cudaErrorInvalidValue;
    }
    *launches += 1;
    return e;
}

cudaError_t perm_dev_end(PermDevice &d, const PermPack &pk, int *launches)
{
    const PermKernelParams &P = d.P;
    const int nk = pk.R * P.ncols;
    if (nk == 0) return cudaSuccess;
    // masked rows are 0 in every column and take part in the maximum
    perm_floor_keys_kernel<<<(nk + 255) / 256, 256, 0, d.st>>>(P.keys, nk, d.pre.any_zero);
    const int eg = (int)std::min<int64_t>((P.V + 127) / 128, (int64_t)d.sm_count * 16);
    if (eg > 0) {
        switch (pk.p) {
#define RMG_EX_CASE(K) case K: perm_exact_voxels_kernel<K><<<eg, 128, 0, d.st>>>(P); break;
            RMG_EX_CASE(1) RMG_EX_CASE(2) RMG_EX_CASE(3) RMG_EX_CASE(4) RMG_EX_CASE(5) RMG_EX_CASE(6) RMG_EX_CASE(7) RMG_EX_CASE(8)
#undef RMG_EX_CASE
        }
    }
    perm_finalize_keys_kernel<<<(nk + 255) / 256, 256, 0, d.st>>>(P.keys, P.pc, P);
    *launches += 3;
    return cudaGetLastError();
}

void perm_dev_release(PermDevice &d)
{
    if (d.cs) {
        cudaStreamSynchronize(d.cs);     // the pinned operand is handed back to the pool after this
        cudaStreamDestroy(d.cs);
        d.cs = nullptr;
    }
    for (auto e : d.ev)
        if (e) cudaEventDestroy(e);
    d.ev.clear();
    if (d.ws) cudaFreeAsync(d.ws, d.st);
    d.ws = nullptr;
}

cudaError_t perm_gemm_run(const PermGemmArgs &a, const PermSet &ps, int sm_count, cudaStream_t st, int *launches)
{
    PermPack pk;
    PermDevice d;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = perm_pack_plan(ps, a.n, a.p, a.hcols, a.ncols, pk);
    if (e == cudaSuccess) e = perm_dev_begin(d, pk, device, sm_count, a, st, nullptr, launches);
    if (e == cudaSuccess) e = perm_dev_prepass(d, a, 0, a.V, launches);     // needs only Y: runs while the host packs
    // batches of permutations: pack on the host (all cores), upload on the side stream, launch the GEMM of the
    // batch; the host packs batch b+1 while the GPU multiplies batch b
    for (int b = 0; b < pk.nbatch && e == cudaSuccess; ++b) {
        perm_pack_batch(ps, pk, b);
        e = perm_dev_upload(d, pk, b);
        if (e == cudaSuccess) e = perm_dev_launch(d, pk, b, b + 1, 0, a.V, launches);
    }
    if (e == cudaSuccess) e = perm_dev_end(d, pk, launches);
    perm_dev_release(d);
    perm_pack_release(pk);
    return e;
}

}  // namespace rmg
