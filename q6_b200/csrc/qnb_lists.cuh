// qnb_lists.cuh -- pair-list generation: cell binning of switch atoms, warp-ballot/scan
// compaction into per-unit neighbour rows, Q-atom partner lists, LRF multipole accumulation.
//
// Replaces nbpplist/nbpwlist/nbwwlist(_box)(_lrf), nbqplist(_box), nbqwlist(_box), cgp_centers and
// lrf_update (nonbondene.f90:43-78, 628-725, 1520-2055, 2622-3085, 3647-4025, 4079-4515).  The
// reference tests all ncgp^2 switch-atom pairs; here switch atoms are binned into cells of edge
// >= the largest cut-off and only the 27 surrounding cells are tested -- with the reference's own
// FP64 expression, so the resulting set of pairs is identical.
//
// LRF kernels: k_lrf_allpairs (sphere with the LRF cut-off spanning it: one thread per target, sources broadcast from
// shared memory) and k_lrf_accumulate (everything else: one warp per target, flat walk over the candidates of the cell
// rows in reach, FP32 screening, accepted items batched through a queue, sources read from k_pack_lrf_planes' records).
#pragma once
#include "qnb_kernels.cuh"

namespace qnb {

// ---------------------------------------------------------------- cells
__device__ __forceinline__ int cell_coord(const Grid &G, double x, int d) {
    int c;
    if (G.periodic) {
        double f = x * G.inv_box[d];
        f -= floor(f);
        c = (int)(f * (double)G.n[d]);
    } else {
        c = (int)floor((x - G.org[d]) * G.inv_cell[d]);
    }
    return min(max(c, 0), G.n[d] - 1);   // clamping is monotone: neighbours stay within +-1 cell
}

// upos[u] = x(switch atom of u); cell_of[u]; histogram
__global__ void k_bin_units(Dev D, Grid G, const double *__restrict__ x, double *__restrict__ upos,
                            int *__restrict__ cell_of, int *__restrict__ cell_count) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= D.nunit) return;
    int a = D.u_sw[u];
    double px = x[3 * a], py = x[3 * a + 1], pz = x[3 * a + 2];
    upos[3 * u] = px; upos[3 * u + 1] = py; upos[3 * u + 2] = pz;
    int c = (cell_coord(G, pz, 2) * G.n[1] + cell_coord(G, py, 1)) * G.n[0] + cell_coord(G, px, 0);
    cell_of[u] = c;
    atomicAdd(&cell_count[c], 1);
}

// single-block exclusive scan: out[i] = sum_{k<i} in[k], out[n] = total
__global__ void k_exclusive_scan(const int *__restrict__ in, int *__restrict__ out, int n) {
    __shared__ int warp_off[32];
    __shared__ int block_total;
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const int v = i < n ? in[i] : 0;
        const int inc = warp_incl_scan(v, lane);
        if (lane == 31) warp_off[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int t = lane < nw ? warp_off[lane] : 0;
            const int ti = warp_incl_scan(t, lane);
            warp_off[lane] = ti - t;
            if (lane == 31) block_total = ti;
        }
        __syncthreads();
        if (i < n) out[i] = carry + warp_off[wid] + inc - v;
        __syncthreads();
        if (tid == 0) carry += block_total;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

__global__ void k_cell_fill(int nunit, const int *__restrict__ cell_of, const int *__restrict__ cell_start,
                            int *__restrict__ cell_cursor, int *__restrict__ cell_items) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nunit) return;
    int c = cell_of[u];
    int p = atomicAdd(&cell_cursor[c], 1);
    cell_items[cell_start[c] + p] = u;
}

// order items inside each cell by unit number so that rows come out in a run-independent order:
// one block per cell, rank of an item = number of smaller unit ids in the same cell
__global__ void k_cell_sort(int ncell, const int *__restrict__ cell_start, const int *__restrict__ unsorted,
                            int *__restrict__ cell_items) {
    const int c = blockIdx.x;
    const int lo = cell_start[c], hi = cell_start[c + 1];
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int v = unsorted[i];
        int rank = 0;
        for (int j = lo; j < hi; j++) rank += unsorted[j] < v;
        cell_items[lo + rank] = v;
    }
}

// cell-ordered copies of what the pair searches read: switch-atom position + unit id, and the number of
// LRF source atoms (non-Q atoms of the unit's charge group)
__global__ void k_pack_items(Dev D, const double *__restrict__ upos, const int *__restrict__ cell_items,
                             double4 *__restrict__ item_pos, float4 *__restrict__ item_posf, float4 *__restrict__ item_scr,
                             int *__restrict__ item_nq) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D.nunit) return;
    const int u = cell_items[idx];
    item_pos[idx] = make_double4(upos[3 * u], upos[3 * u + 1], upos[3 * u + 2], __longlong_as_double((long long)u));
    const int nq = D.u_excl[u] ? 0 : D.g_nq[D.u_grp[u]];
    // screening copy: FP32 position, unit id (-1 - u when the unit has no source atoms: excluded or all Q-atoms)
    double w[3] = {upos[3 * u], upos[3 * u + 1], upos[3 * u + 2]};
    if (D.use_PBC)
        for (int d = 0; d < 3; d++) w[d] -= D.box[d] * floor(w[d] * D.inv_box[d]);   // as binned (cell_coord); screening is min-image
    item_posf[idx] = make_float4((float)w[0], (float)w[1], (float)w[2], __int_as_float(nq > 0 ? u : -1 - u));
    // screening record of the row builder: unit id | entries the unit adds to a partner's row << 24 | excluded << 31
    const int cnt = u < D.ncgp_solute ? D.g_nq[u] : 1;
    item_scr[idx] = make_float4((float)w[0], (float)w[1], (float)w[2],
                                __int_as_float((int)((unsigned)u | ((unsigned)cnt << 24) | (D.u_excl[u] ? 0x80000000u : 0u))));
    item_nq[idx] = nq;
}
// Packed atoms: the non-Q atoms of every non-excluded unit, in cell order of the units.  Row entries are indices
// into this order, so the lanes of a warp (consecutive row entries = neighbouring units of one cell) read
// neighbouring records instead of gathering all over the coordinate array.
// upk[u] = first packed atom of unit u (-1: not packed), pk_sw[p] = switch atom of the packed atom's unit.
__global__ void k_pack_atoms(Dev D, const int *__restrict__ cell_items, const int *__restrict__ src_off,
                             int *__restrict__ pk_atom, float *__restrict__ pk_q, double *__restrict__ pk_qd,
                             int *__restrict__ pk_ct, int *__restrict__ pk_sw, int *__restrict__ upk) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D.nunit) return;
    const int u = cell_items[idx];
    if (D.u_excl[u]) { upk[u] = -1; return; }
    const int g = D.u_grp[u], gf = D.g_first[g], gn = D.g_n[g];
    int p = src_off[idx];
    upk[u] = p;
    const int sw = D.u_sw[u];
    for (int k = 0; k < gn; k++) {
        const int i = D.g_atoms[gf + k];
        if (D.is_q[i]) continue;
        pk_atom[p] = i; pk_q[p] = D.crgf[i]; pk_qd[p] = D.crg[i]; pk_ct[p] = D.ctype[i]; pk_sw[p] = sw;
        p++;
    }
}
// per list build: LRF source records (x,y,z,q) in packed order
__global__ void k_pack_sources(const int *__restrict__ npk, const int *__restrict__ pk_atom, const double *__restrict__ x,
                               const double *__restrict__ crg, double4 *__restrict__ src, float4 *__restrict__ srcf) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= *npk) return;   // number of packed atoms, still on the device when this is launched
    const int i = pk_atom[p];
    const double4 v = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], crg[i]);
    src[p] = v;
    srcf[p] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);   // phi3 path and rsqrt seed of lrf_atom
}
// per list build: the LRF sources once more, laid out for the shell kernel.  Its warps take 32 accepted items at a time, one
// unit per lane, mostly RUNS of consecutive items (cell order): from the per-atom records above a lane's 144 bytes sit 96
// bytes from its neighbour's and one batch costs ~200 L1 wavefronts (r04c: the kernel ran at the same 2.1 ms with 12 or 16
// warps per SM and 12 % fewer instructions - bound by the L1 tag pipe).  Here blocks of 32 items hold PLANES: 12 x 32
// doubles {x,y,z,q of the unit's first three source atoms}, 3 x 32 floats (the charges once more), 32 atom counts
// (kLrfPlaneBytes per block), so the lanes of a run read consecutive words.  Units with fewer than three source atoms are padded with copies
// of the first one carrying no charge (an exact zero contribution); the atoms after the third stay in `src`.  Periodic
// boxes: the unit is moved as a whole by the image that takes its switch atom into the box (the position it is binned and
// screened at), so that the image of the scanned cell row is the pair's shift (k_lrf_accumulate).
constexpr int kLrfPlaneQf = 12 * 32 * 8, kLrfPlaneCnt = kLrfPlaneQf + 3 * 32 * 4, kLrfPlaneBytes = kLrfPlaneCnt + 32 * 4;   // 3584
__global__ void k_pack_lrf_planes(Dev D, int nunit, const double4 *__restrict__ item_pos, const int *__restrict__ src_off,
                                  const double4 *__restrict__ src, char *__restrict__ planes) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nunit) return;
    char *blk = planes + (size_t)(idx >> 5) * kLrfPlaneBytes;
    double *pd = reinterpret_cast<double *>(blk) + (idx & 31);
    float *pf = reinterpret_cast<float *>(blk + kLrfPlaneQf) + (idx & 31);
    const int a0 = src_off[idx], n = src_off[idx + 1] - a0;
    reinterpret_cast<int *>(blk + kLrfPlaneCnt)[idx & 31] = n;
    double sh[3] = {0.0, 0.0, 0.0};
    if (D.use_PBC) {
        const double4 ip = item_pos[idx];
        sh[0] = D.box[0] * floor(ip.x * D.inv_box[0]); sh[1] = D.box[1] * floor(ip.y * D.inv_box[1]); sh[2] = D.box[2] * floor(ip.z * D.inv_box[2]);
    }
    for (int k = 0; k < 3; k++) {
        double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
        if (n > 0) { v = src[a0 + min(k, n - 1)]; if (k >= n) v.w = 0.0; }
        pd[(4 * k + 0) * 32] = v.x - sh[0]; pd[(4 * k + 1) * 32] = v.y - sh[1]; pd[(4 * k + 2) * 32] = v.z - sh[2]; pd[(4 * k + 3) * 32] = v.w;
        pf[k * 32] = (float)v.w;
    }
}
// per step: coordinates in packed order, structure of arrays
__global__ void k_pack_coords(int npk, const int *__restrict__ pk_atom, const double *__restrict__ x,
                              double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npk) return;
    const int i = pk_atom[p];
    px[p] = x[3 * i]; py[p] = x[3 * i + 1]; pz[p] = x[3 * i + 2];
}

// qcp_run: x(atom j) = x_save(atom j) + qcp_coord(j, bead)  (qcp.f90:325-328)
__global__ void k_set_bead(int natq, const int *__restrict__ atoms0, const double *__restrict__ base,
                           const double *__restrict__ disp, double *__restrict__ x) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * natq) return;
    x[3 * atoms0[k / 3] + k % 3] = base[k] + disp[k];
}
// fixed-order sum of the replicated energy slots: dst[k] = sum_s E[s][k]
struct CollectEnergiesBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(int nE, int nslot, const double *__restrict__ E, double *__restrict__ dst, const int BX, const int NBX, const int BY, const int NBY) {
    const int k = BX * blockDim.x + threadIdx.x;
    if (k >= nE) return;
    double e = 0;
    for (int s = 0; s < nslot; s++) e += E[(size_t)s * nE + k];
    dst[k] = e;
}
};
__global__ void
k_collect_energies(int nE, int nslot, const double *__restrict__ E, double *__restrict__ dst) {
    CollectEnergiesBody::run(nE, nslot, E, dst, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// range of cells along one dimension around c with reach m
struct DimRange { int start, count; };
__device__ __forceinline__ DimRange dim_range(int c, int m, int n, int periodic) {
    DimRange r;
    if (!periodic) {
        r.start = max(0, c - m);
        r.count = min(n - 1, c + m) - r.start + 1;
    } else if (2 * m + 1 >= n) {
        r.start = 0; r.count = n;
    } else {
        r.start = c - m; r.count = 2 * m + 1;   // wrapped by the caller
    }
    return r;
}
// up to two contiguous x-segments [lo,hi) of cells
struct XSeg { int lo[2], hi[2], n; };
__device__ __forceinline__ XSeg x_segments(int c, int m, int n, int periodic) {
    XSeg s;
    s.n = 1; s.lo[1] = s.hi[1] = 0;
    if (!periodic) { s.lo[0] = max(0, c - m); s.hi[0] = min(n - 1, c + m) + 1; }
    else if (2 * m + 1 >= n) { s.lo[0] = 0; s.hi[0] = n; }
    else {
        int a = c - m, b = c + m;
        if (a < 0) { s.lo[0] = a + n; s.hi[0] = n; s.lo[1] = 0; s.hi[1] = b + 1; s.n = 2; }
        else if (b >= n) { s.lo[0] = a; s.hi[0] = n; s.lo[1] = 0; s.hi[1] = b - n + 1; s.n = 2; }
        else { s.lo[0] = a; s.hi[0] = b + 1; }
    }
    return s;
}

// ---------------------------------------------------------------- neighbour rows
//
// Row of unit u: [ A_own | A_mir | B ]
//   water  u: A = partner waters (ww),                      B = partner solute ATOMS (pw, forces only)
//   solute u: A = partner solute ATOMS (pp, non-Q, flagged), B = partner waters (pw, owner side)
// A_own holds the pairs the reference lists while looping i = u (owner side), A_mir the rest.
// counts[3*u + {0,1,2}] = |A_own|, |A_mir|, |B|.
template <bool FILL>
__global__ void __launch_bounds__(256)
k_build_rows(Dev D, Cut C, Grid G, int3 reach, const double *__restrict__ x, const double *__restrict__ upos, const int *__restrict__ cell_of,
             const int *__restrict__ cell_start, const double4 *__restrict__ item_pos,
             const int *__restrict__ src_off, int *__restrict__ counts, const int *__restrict__ row_off,
             uint32_t *__restrict__ rows) {
    const int lane = threadIdx.x & 31;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= D.nunit) return;
    const int ns = D.ncgp_solute;
    const bool u_sol = u < ns;
    int n_own = 0, n_mir = 0, n_b = 0;      // warp-uniform cursors
    int base_own = 0, base_mir = 0, base_b = 0;
    if (FILL) {
        base_own = row_off[u];
        base_mir = base_own + counts[3 * u];
        base_b = base_mir + counts[3 * u + 1];
    }
    if (!D.u_excl[u] && !(D.sharded && D.shard_rows && !row_in_any_shard(D, u))) {
        const double pu[3] = {upos[3 * u], upos[3 * u + 1], upos[3 * u + 2]};
        const float puf[3] = {(float)pu[0], (float)pu[1], (float)pu[2]};
        const int cu = cell_of[u];
        const int cx = cu % G.n[0], cy = (cu / G.n[0]) % G.n[1], cz = cu / (G.n[0] * G.n[1]);
        const DimRange rz = dim_range(cz, reach.z, G.n[2], G.periodic), ry = dim_range(cy, reach.y, G.n[1], G.periodic);
        const XSeg xs = x_segments(cx, reach.x, G.n[0], G.periodic);
        const int gs_lo = u_sol ? D.gs_off[u] : 0, gs_hi = u_sol ? D.gs_off[u + 1] : 0;
        for (int iz = 0; iz < rz.count; iz++) {
            int z = rz.start + iz; z = (z % G.n[2] + G.n[2]) % G.n[2];
            for (int iy = 0; iy < ry.count; iy++) {
                int y = ry.start + iy; y = (y % G.n[1] + G.n[1]) % G.n[1];
                const int rowbase = (z * G.n[1] + y) * G.n[0];
                for (int sgi = 0; sgi < xs.n; sgi++) {
                    const int lo = cell_start[rowbase + xs.lo[sgi]], hi = cell_start[rowbase + xs.hi[sgi]];
                    for (int base = lo; base < hi; base += 32) {
                        const int idx = base + lane;
                        double4 ip = make_double4(0, 0, 0, 0);
                        int v = -1;
                        if (idx < hi) { ip = item_pos[idx]; v = (int)__double_as_longlong(ip.w); }
                        bool pass = false, owner_is_u = false;
                        int cls = 0;
                        uint32_t img = 0;   // periodic image of the pair as seen from u (any-atom mode)
                        if (v >= 0 && !D.u_excl[v]) {
                            cls = pair_class(u, v, ns, owner_is_u);
                            if (!(cls == 2 && u == v) && (!D.sharded || (D.shard_rows ? row_in_shard(D, cls, u) : in_shard(D, cls, owner_is_u ? u : v)))) {
                                // FP32 screening first; the FP64 test only decides pairs within ~1e-3 of the cut-off
                                const float r2f = screen_dist2(D, (float)ip.x - puf[0], (float)ip.y - puf[1], (float)ip.z - puf[2]);
                                const int in = (D.any_atom && cls != 2) ? 0 : screen_r2(r2f, (float)C.rc2_of(cls));
                                if (in < 0) pass = true;
                                else if (in == 0) {
                                    const double pv[3] = {ip.x, ip.y, ip.z};
                                    // owner orientation: the reference's outer-loop unit first
                                    const PairTest pt = owner_is_u ? unit_pair_test(D, C, x, cls, u, v, pu, pv)
                                                                   : unit_pair_test(D, C, x, cls, v, u, pv, pu);
                                    pass = pt.listed;
                                    img = (owner_is_u ? pt.img : img_negate(pt.img)) << kImgShift;
                                }
                            }
                        }
                        // entries emitted by this lane, by segment
                        const bool v_sol = v >= 0 && v < ns;
                        int c_own = 0, c_mir = 0, c_b = 0;
                        if (pass) {
                            if (u_sol) {
                                if (v_sol) { if (owner_is_u) c_own = D.g_nq[v]; else c_mir = D.g_nq[v]; }
                                else c_b = 1;
                            } else {
                                if (v_sol) c_b = D.g_nq[v];
                                else { if (owner_is_u) c_own = 1; else c_mir = 1; }
                            }
                        }
                        const int s_own = warp_incl_scan(c_own, lane), s_mir = warp_incl_scan(c_mir, lane),
                                  s_b = warp_incl_scan(c_b, lane);
                        if (FILL && pass) {
                            int p_own = base_own + n_own + s_own - c_own;
                            int p_mir = base_mir + n_mir + s_mir - c_mir;
                            int p_b = base_b + n_b + s_b - c_b;
                            const uint32_t pk0 = (uint32_t)src_off[idx];   // first packed atom of unit v
                            if (v_sol) {
                                // flatten the group's non-Q atoms (entries = packed atom indices)
                                const int gf = D.g_first[v], gn = D.g_n[v];
                                uint32_t flag = ((u_sol && owner_is_u) ? kOwnerBit : 0u) | img;
                                int p = u_sol ? (owner_is_u ? p_own : p_mir) : p_b;
                                uint32_t pk = pk0;
                                for (int k = 0; k < gn; k++) {
                                    int b = D.g_atoms[gf + k];
                                    if (D.is_q[b]) continue;
                                    uint32_t e = (pk++) | flag;
                                    if (u_sol) {
                                        // special if b relates to any atom of group u (excluded / 1-4 / same group)
                                        int l = gs_lo, h = gs_hi;
                                        while (l < h) { int m = (l + h) >> 1; if (D.gs_atoms[m] < b) l = m + 1; else h = m; }
                                        if (l < gs_hi && D.gs_atoms[l] == b) e |= kSpecialBit;
                                    }
                                    rows[p++] = e;
                                }
                            } else {
                                // a water partner is named by the packed index of its first atom (O)
                                if (u_sol) rows[p_b] = pk0 | kOwnerBit | img;
                                else if (owner_is_u) rows[p_own] = pk0 | kOwnerBit;
                                else rows[p_mir] = pk0;
                            }
                        }
                        n_own += __shfl_sync(kFull, s_own, 31);
                        n_mir += __shfl_sync(kFull, s_mir, 31);
                        n_b += __shfl_sync(kFull, s_b, 31);
                    }
                }
            }
        }
    }
    if (!FILL && lane == 0) {
        counts[3 * u] = n_own; counts[3 * u + 1] = n_mir; counts[3 * u + 2] = n_b;
    }
}

// ---------------------------------------------------------------- neighbour rows, flattened scan (round 2)
// Same rows as k_build_rows, built for issue efficiency (r02e: 302 M instructions per pass on the 98 k-atom box, 280 per
// 32-candidate step, a third of the lanes idle because every step took ONE cell-row segment of ~21 candidates, two
// dependent cell_start loads and integer divisions per step, FP64 positions converted lane by lane):
//  * the (<= 2 x rows) item ranges of the scanned cells are tabulated once per unit in shared memory, all cell_start
//    loads in parallel, and the candidates of all ranges are walked as ONE sequence in full steps of 32;
//  * a 16-byte FP32 screening record per candidate (position, unit, entry count, excluded flag) decides all but the
//    borderline candidates (|r2 - Rc2| < 1e-3 Rc2 + 0.05), which take the reference's FP64 expression as before;
//  * positions in the segments come from ballots when no solute group is among the accepted candidates.
constexpr int kRowSegCap = 128;    // item ranges tabulated per batch (5 x 5 rows x 2 x-segments = 50 with reach 2)
constexpr int kRowScanWarps = 8;
template <bool FILL>
__global__ void __launch_bounds__(32 * kRowScanWarps)
k_rows_scan(Dev D, Cut C, Grid G, int3 reach, const double *__restrict__ x, const double *__restrict__ upos, const int *__restrict__ cell_of,
            const int *__restrict__ cell_start, const double4 *__restrict__ item_pos, const float4 *__restrict__ item_scr,
            const int *__restrict__ src_off, int *__restrict__ counts, const int *__restrict__ row_off,
            uint32_t *__restrict__ rows) {
    __shared__ int s_lo[kRowScanWarps][kRowSegCap], s_cum[kRowScanWarps][kRowSegCap + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= D.nunit) return;
    const int ns = D.ncgp_solute;
    const bool u_sol = u < ns;
    int n_own = 0, n_mir = 0, n_b = 0;      // warp-uniform cursors
    int base_own = 0, base_mir = 0, base_b = 0;
    if (FILL) {
        base_own = row_off[u];
        base_mir = base_own + counts[3 * u];
        base_b = base_mir + counts[3 * u + 1];
    }
    if (!D.u_excl[u] && !(D.sharded && D.shard_rows && !row_in_any_shard(D, u))) {
        const double pu[3] = {upos[3 * u], upos[3 * u + 1], upos[3 * u + 2]};
        float puf[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double w = pu[d];
            if (D.use_PBC) w -= D.box[d] * floor(w * D.inv_box[d]);   // as the records are stored (k_pack_items)
            puf[d] = f32_once(w);
        }
        const float bx = f32_once(D.box[0]), by = f32_once(D.box[1]), bz = f32_once(D.box[2]);
        const float ibx = f32_once(D.inv_box[0]), iby = f32_once(D.inv_box[1]), ibz = f32_once(D.inv_box[2]);
        // screening thresholds by the partner's kind (see screen_r2)
        const float rc_s = f32_once(C.rc2_of(u_sol ? 0 : 1)), rc_w = f32_once(C.rc2_of(u_sol ? 1 : 2));
        const float lo_s = rc_s * (1.0f - 1e-3f) - 0.05f, hi_s = rc_s * (1.0f + 1e-3f) + 0.05f;
        const float lo_w = rc_w * (1.0f - 1e-3f) - 0.05f, hi_w = rc_w * (1.0f + 1e-3f) + 0.05f;
        const int cu = cell_of[u];
        const int cx = cu % G.n[0], cy = (cu / G.n[0]) % G.n[1], cz = cu / (G.n[0] * G.n[1]);
        const DimRange rz = dim_range(cz, reach.z, G.n[2], G.periodic), ry = dim_range(cy, reach.y, G.n[1], G.periodic);
        const XSeg xs = x_segments(cx, reach.x, G.n[0], G.periodic);
        const int xlo0 = xs.lo[0], xhi0 = xs.hi[0], xlo1 = xs.lo[1], xhi1 = xs.hi[1];
        const int gs_lo = u_sol ? D.gs_off[u] : 0, gs_hi = u_sol ? D.gs_off[u + 1] : 0;
        const int nseg = rz.count * ry.count * xs.n;
        const int a_u = u_sol ? u + 1 : u - ns + 1;   // 1-based number inside its kind (checkerboard rule)
        for (int t0 = 0; t0 < nseg; t0 += kRowSegCap) {
            const int nb = min(kRowSegCap, nseg - t0);
            // ---- item ranges of this batch of cell-row segments, then their running offsets
            __syncwarp();
            int carry = 0;
            for (int tb = 0; tb < nb; tb += 32) {
                const int t = t0 + tb + lane;
                int lo = 0, len = 0;
                if (tb + lane < nb) {
                    const int sgi = xs.n == 2 ? (t & 1) : 0, r = xs.n == 2 ? (t >> 1) : t;
                    const int iz = r / ry.count, iy = r - iz * ry.count;
                    int z = rz.start + iz, y = ry.start + iy;
                    z += z < 0 ? G.n[2] : 0; z -= z >= G.n[2] ? G.n[2] : 0;
                    y += y < 0 ? G.n[1] : 0; y -= y >= G.n[1] ? G.n[1] : 0;
                    const int rowbase = (z * G.n[1] + y) * G.n[0];
                    lo = cell_start[rowbase + (sgi ? xlo1 : xlo0)];
                    len = cell_start[rowbase + (sgi ? xhi1 : xhi0)] - lo;
                    s_lo[wib][tb + lane] = lo;
                }
                const int inc = warp_incl_scan(len, lane);
                if (tb + lane < nb) s_cum[wib][tb + lane] = carry + inc - len;
                carry += __shfl_sync(kFull, inc, 31);
            }
            if (lane == 0) s_cum[wib][nb] = carry;
            __syncwarp();
            const int total = carry;
            int s0 = 0;
            for (int v0 = 0; v0 < total; v0 += 32) {
                const int vi = v0 + lane;
                const bool act = vi < total;
                int sg = s0;
                if (act) while (vi >= s_cum[wib][sg + 1]) sg++;
                s0 = __shfl_sync(kFull, sg, 31);
                const int idx = act ? s_lo[wib][sg] + (vi - s_cum[wib][sg]) : 0;
                bool pass = false, owner_is_u = false;
                int v = -1, cnt = 0;
                uint32_t img = 0;   // periodic image of the pair as seen from u (any-atom mode)
                if (act) {
                    const float4 pf = item_scr[idx];
                    const uint32_t sb = (uint32_t)__float_as_int(pf.w);
                    v = (int)(sb & kIdMask);
                    cnt = (int)((sb >> 24) & 0x7fu);
                    const bool v_sol = v < ns;
                    bool consider = !(sb & 0x80000000u) && !(!u_sol && u == v);
                    // class and owner side (pair_class): pw is owned by the solute group, else the checkerboard rule
                    int cls;
                    if (u_sol != v_sol) { cls = 1; owner_is_u = u_sol; }
                    else {
                        cls = u_sol ? 0 : 2;
                        const int a_v = v_sol ? v + 1 : v - ns + 1;
                        const bool even = ((a_u + a_v) & 1) == 0;
                        owner_is_u = a_u == a_v ? true : (a_u > a_v ? !even : even);
                    }
                    if (consider && D.sharded) consider = D.shard_rows ? row_in_shard(D, cls, u) : in_shard(D, cls, owner_is_u ? u : v);
                    if (consider) {
                        float dx = pf.x - puf[0], dy = pf.y - puf[1], dz = pf.z - puf[2];
                        if (D.use_PBC) { dx -= bx * rintf(dx * ibx); dy -= by * rintf(dy * iby); dz -= bz * rintf(dz * ibz); }
                        const float r2f = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        const float lo = v_sol ? lo_s : lo_w, hi = v_sol ? hi_s : hi_w;
                        if (!(D.any_atom && cls != 2) && r2f < lo) pass = true;
                        else if ((D.any_atom && cls != 2) || !(r2f > hi)) {
                            const double4 ip = item_pos[idx];
                            const double pv[3] = {ip.x, ip.y, ip.z};
                            // owner orientation: the reference's outer-loop unit first
                            const PairTest pt = owner_is_u ? unit_pair_test(D, C, x, cls, u, v, pu, pv)
                                                           : unit_pair_test(D, C, x, cls, v, u, pv, pu);
                            pass = pt.listed;
                            img = (owner_is_u ? pt.img : img_negate(pt.img)) << kImgShift;
                        }
                    }
                }
                const bool v_sol = v >= 0 && v < ns;
                // segment of the entries this lane emits: 0 own, 1 mirror, 2 other kind
                const int seg = (u_sol != v_sol) ? 2 : (owner_is_u ? 0 : 1);
                const int c_emit = pass ? cnt : 0;
                int p_own, p_mir, p_b;   // exclusive positions inside this step
                int t_own, t_mir, t_b;   // totals of this step
                if (!__any_sync(kFull, pass && v_sol)) {
                    // one entry per accepted candidate: ballots
                    const unsigned lt = (1u << lane) - 1u;
                    const unsigned m0 = __ballot_sync(kFull, pass && seg == 0), m1 = __ballot_sync(kFull, pass && seg == 1),
                                   m2 = __ballot_sync(kFull, pass && seg == 2);
                    p_own = __popc(m0 & lt); p_mir = __popc(m1 & lt); p_b = __popc(m2 & lt);
                    t_own = __popc(m0); t_mir = __popc(m1); t_b = __popc(m2);
                } else {
                    const int c0 = seg == 0 ? c_emit : 0, c1 = seg == 1 ? c_emit : 0, c2 = seg == 2 ? c_emit : 0;
                    const int s_own = warp_incl_scan(c0, lane), s_mir = warp_incl_scan(c1, lane), s_b = warp_incl_scan(c2, lane);
                    p_own = s_own - c0; p_mir = s_mir - c1; p_b = s_b - c2;
                    t_own = __shfl_sync(kFull, s_own, 31); t_mir = __shfl_sync(kFull, s_mir, 31); t_b = __shfl_sync(kFull, s_b, 31);
                }
                if (FILL && pass) {
                    const uint32_t pk0 = (uint32_t)src_off[idx];   // first packed atom of unit v
                    if (v_sol) {
                        // flatten the group's non-Q atoms (entries = packed atom indices)
                        const int gf = D.g_first[v], gn = D.g_n[v];
                        const uint32_t flag = ((u_sol && owner_is_u) ? kOwnerBit : 0u) | img;
                        int p = u_sol ? (owner_is_u ? base_own + n_own + p_own : base_mir + n_mir + p_mir) : base_b + n_b + p_b;
                        uint32_t pk = pk0;
                        for (int k = 0; k < gn; k++) {
                            const int b = D.g_atoms[gf + k];
                            if (D.is_q[b]) continue;
                            uint32_t e = (pk++) | flag;
                            if (u_sol) {
                                // special if b relates to any atom of group u (excluded / 1-4 / same group)
                                int l = gs_lo, h = gs_hi;
                                while (l < h) { const int m = (l + h) >> 1; if (D.gs_atoms[m] < b) l = m + 1; else h = m; }
                                if (l < gs_hi && D.gs_atoms[l] == b) e |= kSpecialBit;
                            }
                            rows[p++] = e;
                        }
                    } else {
                        // a water partner is named by the packed index of its first atom (O)
                        if (u_sol) rows[base_b + n_b + p_b] = pk0 | kOwnerBit | img;
                        else if (owner_is_u) rows[base_own + n_own + p_own] = pk0 | kOwnerBit;
                        else rows[base_mir + n_mir + p_mir] = pk0;
                    }
                }
                n_own += t_own; n_mir += t_mir; n_b += t_b;
            }
        }
    }
    if (!FILL && lane == 0) {
        counts[3 * u] = n_own; counts[3 * u + 1] = n_mir; counts[3 * u + 2] = n_b;
    }
}

// ---- chunked re-layout of the rows of units [u0, u0+n): per unit and per i-tile ceil((nown+nmir)/32) chunks of
// kind A and ceil(nb/32) chunks of kind B.  tile_atoms = 0: one tile per unit (waters); else tiles of that many
// non-Q atoms of the unit's charge group (solute).
__device__ __forceinline__ int unit_tiles(const Dev &D, int u, int tile_atoms) {
    return tile_atoms == 0 ? 1 : (D.g_nq[u] + tile_atoms - 1) / tile_atoms;
}
// Cost of a chunk and of a tile change (load of the own atoms + flush of their gradient) for the static work split of
// the persistent force kernels, in units of 10 cycles: least-squares fit of the per-warp busy time against the per-warp
// counts on C2 and C5 (tools/exp_trace.py, -DQNB_TRACE).  own = carries FP64 energies, mir = forces only, b = the
// other unit kind.
struct ChunkCost { int own, mir, b, tile; };
// new_rows: the round-2 gradient-only kernels (qnb_rows.cuh), where own and mirror chunks cost the same (issue slots per
// chunk counted in the SASS: water A 105, B 70; solute A 130, B 290; tile change ~60 / ~90)
__device__ __forceinline__ ChunkCost chunk_cost(int tile_atoms, bool new_rows) {
    if (new_rows) return tile_atoms == 0 ? ChunkCost{105, 105, 70, 60} : ChunkCost{130, 130, 290, 90};
    return tile_atoms == 0 ? ChunkCost{110, 75, 68, 80} : ChunkCost{165, 145, 510, 160};
}
// chunk counts of one unit's tile: own-carrying A chunks, mirror-only A chunks, B chunks (own entries come first)
__device__ __forceinline__ void unit_chunks(const int *__restrict__ counts, int u, int &n_own, int &n_mir, int &n_b) {
    const int own = counts[3 * u], na = own + counts[3 * u + 1];
    n_own = (own + 31) / 32;
    n_mir = (na + 31) / 32 - n_own;
    n_b = (counts[3 * u + 2] + 31) / 32;
}
__global__ void k_chunk_count(Dev D, int u0, int n, int tile_atoms, bool new_rows, const int *__restrict__ counts, int *__restrict__ nch,
                              int *__restrict__ ucost) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int u = u0 + k;
    int n_own, n_mir, n_b;
    unit_chunks(counts, u, n_own, n_mir, n_b);
    const ChunkCost w = chunk_cost(tile_atoms, new_rows);
    const int tiles = unit_tiles(D, u, tile_atoms);
    nch[k] = tiles * (n_own + n_mir + n_b);
    ucost[k] = (n_own + n_mir + n_b) > 0 ? tiles * (n_own * w.own + n_mir * w.mir + n_b * w.b + w.tile) : 0;
}
// first chunk of every warp of a persistent force kernel: equal shares of the summed chunk costs.
// wstart[w] = smallest chunk c whose cost prefix reaches w*total/nwarp; wstart[nwarp] = number of chunks.
__global__ void k_warp_starts(Dev D, int u0, int n, int tile_atoms, bool new_rows, const int *__restrict__ counts,
                              const int *__restrict__ choff, const int *__restrict__ cost_off, int nwarp,
                              int *__restrict__ wstart) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > nwarp) return;
    if (w == nwarp) { wstart[w] = choff[n]; return; }
    const long long total = cost_off[n];
    const int target = (int)((total * w) / nwarp);
    int lo = 0, hi = n;   // last unit k with cost_off[k] <= target
    while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (cost_off[m] <= target) lo = m; else hi = m; }
    const int k = lo, u = u0 + k;
    int n_own, n_mir, n_b;
    unit_chunks(counts, u, n_own, n_mir, n_b);
    const ChunkCost cw = chunk_cost(tile_atoms, new_rows);
    const int tiles = unit_tiles(D, u, tile_atoms);
    int c = choff[k], acc = cost_off[k];
    for (int tile = 0; tile < tiles && acc < target && n_own + n_mir + n_b > 0; tile++) {
        acc += cw.tile;
        for (int seg = 0; seg < 3 && acc < target; seg++) {
            const int m = seg == 0 ? n_own : seg == 1 ? n_mir : n_b, cst = seg == 0 ? cw.own : seg == 1 ? cw.mir : cw.b;
            for (int j = 0; j < m && acc < target; j++) { acc += cst; c++; }
        }
        // the walk stopped inside this tile, or the tile is complete and c stands at the next one
    }
    wstart[w] = min(c, choff[n]);
}
// one warp per unit: copy the segments into 32-entry chunks padded with 0xffffffff; descriptor = {unit - u0, kind | tile << 8}
__global__ void k_chunk_fill(Dev D, int u0, int n, int tile_atoms, const int *__restrict__ counts,
                             const int *__restrict__ row_off, const uint32_t *__restrict__ rows,
                             const int *__restrict__ choff, int2 *__restrict__ cdesc, uint32_t *__restrict__ crow,
                             const int *__restrict__ pk_atom, uint16_t *__restrict__ cspec, bool mark_kind, uint32_t pad_id) {
    const int lane = threadIdx.x & 31;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n) return;
    const int u = u0 + k;
    int c = choff[k];
    const int ntile = unit_tiles(D, u, tile_atoms);
    const int na = counts[3 * u] + counts[3 * u + 1], nb = counts[3 * u + 2];
    for (int tile = 0; tile < ntile; tile++) {
        int base = row_off[u];
        for (int seg = 0; seg < 2; seg++) {
            const int m = seg == 0 ? na : nb;
            for (int b = 0; b < m; b += 32, c++) {
                if (lane == 0) cdesc[c] = make_int2(k, seg | (tile << 8));
                uint32_t e = (b + lane < m) ? rows[base + b + lane] : 0xffffffffu;
                // water rows of the pipelined kernel (k_water_rows): bit 30 of EVERY lane names the chunk's kind, so the
                // loads of a chunk can be issued from its entries alone; a padding lane names the all-zero record kept
                // behind the last packed atom (pad_id), so that it needs no test before the load
                if (mark_kind) {
                    if (e == 0xffffffffu) e = pad_id;
                    e = seg == 0 ? (e & ~kSpecialBit) : (e | kSpecialBit);
                }
                crow[(size_t)c * 32 + lane] = e;
                if (cspec) {
                    // solute tiles: resolve the special pairs (exclusion lists, 1-4 neighbours, own group) of this
                    // partner against every tile atom now, 3 bits each: 1 skip, 2 1-4 pair, 4 energy on the other side
                    unsigned sp = 0;
                    if (seg == 0 && e != 0xffffffffu && (e & kSpecialBit)) {
                        const int bb = pk_atom[e & kIdMask];
                        const bool same = D.grp_of_atom[bb] == u;
                        const int k0 = D.nq_off[u] + tile * tile_atoms, nt = min(tile_atoms, D.nq_off[u + 1] - k0);
                        for (int t = 0; t < nt; t++) {
                            const int a = D.nq_atoms[k0 + t];
                            unsigned bits = 0;
                            if (a == bb) bits = 1;
                            else {
                                const int sc = special_code(D, a, bb);
                                if (sc == 0) bits = 1;
                                else if (sc == 3) bits = 2;
                            }
                            if (same && !(a < bb)) bits |= 4;
                            sp |= bits << (3 * t);
                        }
                    }
                    cspec[(size_t)c * 32 + lane] = (uint16_t)sp;
                }
            }
            base += m;
        }
    }
}

__global__ void k_row_totals(int nunit, const int *__restrict__ counts, int *__restrict__ tot) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < nunit) tot[u] = counts[3 * u] + counts[3 * u + 1] + counts[3 * u + 2];
}

// ---------------------------------------------------------------- Q-atom partner lists
// nbqplist / nbqplist_box (L3647-3859): flag solute atoms that pair with every Q-atom; for the box the
// group's reference atom for the periodic shift (first atom found inside Rcq) is kept per group.
__global__ void k_qp_flags(Dev D, Cut C, const double *__restrict__ x, int *__restrict__ flag,
                           int *__restrict__ qp_shift_atom) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.ncgp_solute) return;
    const int gf = D.g_first[g], gn = D.g_n[g];
    bool inside = false;
    int ref = -1;
    if (g + 1 >= D.qp_s && g + 1 <= D.qp_e) {
        if (!D.use_PBC) {
            int ia = D.g_switch[g];
            if (!D.excl[ia]) {
                if (!D.any_atom) {
                    // nbqplist: r2 = q_dist4(x(ia),xpcent); skip if r2 > rcut2 (L3707-3710)
                    double r2 = sq3(__dsub_rn(D.xpcent[0], x[3 * ia]), __dsub_rn(D.xpcent[1], x[3 * ia + 1]),
                                    __dsub_rn(D.xpcent[2], x[3 * ia + 2]));
                    inside = !(r2 > C.rcq2);
                } else {
                    // nbqplis2: any atom of the group within Rcq of xpcent (L3466-3476)
                    for (int k = 0; k < gn && !inside; k++) {
                        const int i = D.g_atoms[gf + k];
                        double r2 = sq3(__dsub_rn(D.xpcent[0], x[3 * i]), __dsub_rn(D.xpcent[1], x[3 * i + 1]),
                                        __dsub_rn(D.xpcent[2], x[3 * i + 2]));
                        inside = r2 <= C.rcq2;
                    }
                }
            }
        } else if (C.Rq < 0.0) {
            inside = true;   // no cut-off: every group, no registered reference atom (zero shift)
        } else {
            const int qs = D.qswitch0;
            for (int k = 0; k < gn && !inside; k++) {
                int i = D.g_atoms[gf + k];
                double sx = __dsub_rn(x[3 * i], x[3 * qs]), sy = __dsub_rn(x[3 * i + 1], x[3 * qs + 1]),
                       sz = __dsub_rn(x[3 * i + 2], x[3 * qs + 2]);
                double r2 = sq3(__dsub_rn(pshift(sx, D.box[0], D.inv_box[0]), sx),
                                __dsub_rn(pshift(sy, D.box[1], D.inv_box[1]), sy),
                                __dsub_rn(pshift(sz, D.box[2], D.inv_box[2]), sz));
                if (r2 <= C.rcq2) { inside = true; ref = i; }
            }
        }
    }
    qp_shift_atom[g] = ref;
    for (int k = 0; k < gn; k++) {
        int i = D.g_atoms[gf + k];
        // "check if already on qq list": any(qconn(:,i,:) <= 3) also covers the Q-atoms themselves
        flag[i] = (inside && i < D.nat_solute && !D.qbonded[i]) ? 1 : 0;
    }
}
// nbqwlist / nbqwlist_box (L3862-4025)
__global__ void k_qw_flags(Dev D, Cut C, const double *__restrict__ x, int *__restrict__ flag) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= D.nwat) return;
    int ia = D.nat_solute + 3 * w;
    bool inside = false;
    if (w + 1 >= D.qw_s && w + 1 <= D.qw_e) {
        if (!D.use_PBC) {
            if (!D.excl[ia]) {
                double r2 = sq3(__dsub_rn(D.xpcent[0], x[3 * ia]), __dsub_rn(D.xpcent[1], x[3 * ia + 1]),
                                __dsub_rn(D.xpcent[2], x[3 * ia + 2]));
                inside = r2 <= C.rcq2;
            }
        } else if (C.Rq > 0.0) {
            const int qs = D.qswitch0;
            double sx = __dsub_rn(x[3 * ia], x[3 * qs]), sy = __dsub_rn(x[3 * ia + 1], x[3 * qs + 1]),
                   sz = __dsub_rn(x[3 * ia + 2], x[3 * qs + 2]);
            double r2 = sq3(__dsub_rn(pshift(sx, D.box[0], D.inv_box[0]), sx),
                            __dsub_rn(pshift(sy, D.box[1], D.inv_box[1]), sy),
                            __dsub_rn(pshift(sz, D.box[2], D.inv_box[2]), sz));
            inside = (r2 <= C.rcq2);
        } else {
            // Rq <= 0: r2 = one; listed if (r2 <= rcut2) .or. (Rq < zero) (L3996-4003)
            inside = (1.0 <= C.rcq2) || (C.Rq < 0.0);
        }
    }
    flag[w] = inside ? 1 : 0;
}
__global__ void k_compact(int n, const int *__restrict__ flag, const int *__restrict__ pos, int *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = i;
}

// ---------------------------------------------------------------- LRF
// lrf layout per group: 43 doubles = cgp_cent(3) phi0 phi1(3) phi2(3x3) phi3(9x3)  (LRF_TYPE, globals.f90:503)

// cgp_centers (nonbondene.f90:43-78)
__global__ void k_cgp_centers(Dev D, const double *__restrict__ x, double *__restrict__ lrf) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.ncgp) return;
    double cx = 0, cy = 0, cz = 0;
    const int gf = D.g_first[g], gn = D.g_n[g];
    for (int k = 0; k < gn; k++) {
        int a = D.g_atoms[gf + k];
        cx = __dadd_rn(cx, x[3 * a]); cy = __dadd_rn(cy, x[3 * a + 1]); cz = __dadd_rn(cz, x[3 * a + 2]);
    }
    double *l = lrf + (size_t)QNB_LRF_STRIDE * g;
    const double n = (double)gn;
    l[0] = cx / n; l[1] = cy / n; l[2] = cz / n;
    for (int k = 3; k < QNB_LRF_STRIDE; k++) l[k] = 0.0;
}

// centres only (after the multi-GPU sum of the moments, lrf_add keeps cgp_cent: nonbondene.f90:582)
__global__ void k_cgp_centers_only(Dev D, const double *__restrict__ x, double *__restrict__ lrf) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.ncgp) return;
    double cx = 0, cy = 0, cz = 0;
    const int gf = D.g_first[g], gn = D.g_n[g];
    for (int k = 0; k < gn; k++) {
        int a = D.g_atoms[gf + k];
        cx = __dadd_rn(cx, x[3 * a]); cy = __dadd_rn(cy, x[3 * a + 1]); cz = __dadd_rn(cz, x[3 * a + 2]);
    }
    double *l = lrf + (size_t)QNB_LRF_STRIDE * g;
    const double n = (double)gn;
    l[0] = cx / n; l[1] = cy / n; l[2] = cz / n;
}

constexpr int kLrfSegBatch = 128;
__constant__ int kLrfExpand[40] = {0, 1, 2, 3,
                                   4, 5, 6, 5, 7, 8, 6, 8, 9,
                                   10, 11, 12, 11, 13, 14, 12, 14, 15,    // n = x: (xx*, xy*, xz*)
                                   11, 13, 14, 13, 16, 17, 14, 17, 18,    // n = y
                                   12, 14, 15, 14, 17, 18, 15, 18, 19};   // n = z

// ---- lrf_update of one source atom (nonbondene.f90:656-719) in RAW moments: the traceless combinations the reference
// adds atom by atom (phi2 = f1 dr dr - f0 I, phi3 = 5 f2 dr dr dr - f2 r^2 (...)) are linear in the sums
//   m[0] = S q/r            m[1..3] = S f0 dr          m[4..9] = S (f0/r^2) dr dr (xx xy xz yy yz zz)   m[10] = S f0
//   h[0..9] = S (q/r^7) dr dr dr (xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz)                h[10..12] = S (q/r^5) dr
// with f0 = q/r^3, so the atom loop is 29 FP64 + 34 FP32 operations instead of 32 + 49 and lrf_finish() forms
// phi0, phi1 = -m[1..3], phi2 = 3 m[4..9] - m[10] I, phi3 = -15 h[0..9] + {9, 3} h[10..12] once per target.
// phi0-phi2 FP64 (FP32 phi2 was measured to move E%LRF by 1.1e-6 relative); phi3 enters only the field, as
// 1/2 dr.phi3.dr with |dr| ~ 1 A against r >= Rc: a (dr/r)^2 ~ 1e-2 correction to phi1, summed in FP32.
// d: displacement in FP64, f: the same in FP32 (phi3 path and rsqrt seed; its 1e-6 error is cubed away by the Halley step)
__device__ __forceinline__ void lrf_atom(double (&m)[11], float (&h)[13], double dx, double dy, double dz, double q,
                                         float fx, float fy, float fz, float qf) {
    const double r2 = dx * dx + dy * dy + dz * dz;
    const float rif = rsqrt_fast(fmaf(fx, fx, fmaf(fy, fy, fz * fz)));
    const double ri = rsqrt_refine(r2, rif), ri2 = ri * ri;
    const double qri = q * ri, f0 = qri * ri2, f1 = f0 * ri2;
    m[0] += qri;
    m[10] += f0;
    m[1] += dx * f0; m[2] += dy * f0; m[3] += dz * f0;
    const double tx = f1 * dx, ty = f1 * dy, tz = f1 * dz;
    m[4] += tx * dx; m[5] += tx * dy; m[6] += tx * dz; m[7] += ty * dy; m[8] += ty * dz; m[9] += tz * dz;
    const float rif2 = rif * rif, p5 = qf * rif * rif2 * rif2, p7 = p5 * rif2;
    h[10] += p5 * fx; h[11] += p5 * fy; h[12] += p5 * fz;
    const float ax = p7 * fx, ay = p7 * fy, az = p7 * fz;
    const float axx = ax * fx, axy = ax * fy, axz = ax * fz, ayy = ay * fy, ayz = ay * fz, azz = az * fz;
    h[0] += axx * fx; h[1] += axx * fy; h[2] += axx * fz; h[3] += axy * fy; h[4] += axy * fz; h[5] += axz * fz;
    h[6] += ayy * fy; h[7] += ayy * fz; h[8] += ayz * fz; h[9] += azz * fz;
}
// raw sums (24 values: m[0..10], h[0..12] as doubles) -> the 20 unique moments in kLrfExpand's order
__device__ __forceinline__ double lrf_finish(const double *raw, int k) {
    const double *m = raw, *h = raw + 11;
    if (k == 0) return m[0];
    if (k <= 3) return -m[k];
    if (k <= 9) return 3.0 * m[k] - ((k == 4 || k == 7 || k == 9) ? m[10] : 0.0);
    // phi3, unique index k-10: xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz; trace part -(-3 q/r^5) (...): 3x for aaa, 1x for aab
    const int u = k - 10;
    const double t = -15.0 * h[u];
    switch (u) {
    case 0: return t + 9.0 * h[10];   // xxx
    case 1: return t + 3.0 * h[11];   // xxy
    case 2: return t + 3.0 * h[12];   // xxz
    case 3: return t + 3.0 * h[10];   // xyy
    case 4: return t;                 // xyz
    case 5: return t + 3.0 * h[10];   // xzz
    case 6: return t + 9.0 * h[11];   // yyy
    case 7: return t + 3.0 * h[12];   // yyz
    case 8: return t + 3.0 * h[11];   // yzz
    default: return t + 9.0 * h[12];  // zzz
    }
}
constexpr int kLrfRaw = 24;
// the same with the FP32 displacement rounded from the FP64 one (plane records carry no FP32 coordinates)
__device__ __forceinline__ void lrf_atom_d(double (&m)[11], float (&h)[13], double dx, double dy, double dz, double q, float qf) {
    lrf_atom(m, h, dx, dy, dz, q, (float)dx, (float)dy, (float)dz, qf);
}

// lrf_update (nonbondene.f90:628-725), gathered per TARGET group: the warp of target unit t sums the
// contribution of every source atom whose unit pair (t,s) the reference sends through the LRF branch
// (outside the class cut-off, inside RcLRF, pair owned by this shard).  20 unique moments are
// accumulated (phi2 and phi3 are symmetric) and expanded on write.  FP64, no divisions:
// field0 = q/r^3, field1 = 3 field0/r^2, field2 = -field1/r^2 from 1/r (rsqrt seed + Halley step).
// Scan modes.  ROWSHIFT (periodic box, reach small enough that the image of a scanned cell row is known from its cell
// offset, 2(m+2) <= n in every dimension, see qnb.cu): the periodic image is applied once per cell row to the TARGET, rows that
// cannot reach the LRF shell are skipped and the x-range of the others is trimmed to the shell's chord, so the
// per-candidate screening is nine FP32 instructions.  GENERAL adds what any-atom cut-offs and sharded builds need.
struct LrfSeg { int lo, hi; float tx, ty, tz; int img; };   // img: periodic image of the row, (ix+1) | (iy+1) << 2 | (iz+1) << 4
#ifndef QNB_LRF_MINB
#define QNB_LRF_MINB 4
#endif
#ifndef QNB_LRF_DEPTH
#define QNB_LRF_DEPTH 3   // register sets of the screening walk: records of DEPTH-1 steps in flight
#endif
template <bool COMPACT, bool ROWSHIFT, bool GENERAL>
__global__ void __launch_bounds__(32 * kRowWarps, QNB_LRF_MINB)
k_lrf_accumulate(Dev D, Cut C, Grid G, int3 reach, const double *__restrict__ x, const double *__restrict__ upos,
                 const int *__restrict__ cell_of, const int *__restrict__ cell_start, const int *__restrict__ cell_items,
                 const double4 *__restrict__ item_pos, const float4 *__restrict__ item_posf,
                 const int *__restrict__ src_off, const double4 *__restrict__ src, const float4 *__restrict__ srcf,
                 const char *__restrict__ planes, double *__restrict__ lrf) {
    // One WARP per target unit, the kRowWarps targets of a block being neighbours in cell order: they scan (nearly) the
    // same cells at the same time, so the source records one warp pulls through L1 serve the others (r02e: one block per
    // target with its warps on different cell rows, 49 % L1 hit rate, 258 KB of source records per target through L2).
    // Candidates are first screened (FP32 distance with a safety band, exact FP64 test inside the band) and the
    // accepted ones compacted into a per-warp queue, so that the expensive accumulation always runs on full warps
    // even when only a few percent of the scanned cells' units lie inside the LRF shell (periodic boxes).
    // everything a warp keeps in shared memory behind ONE base address (r04e: five separate per-warp arrays had their
    // addresses rebuilt from threadIdx in every step, 39 of ~110 instructions per step)
    struct WarpSm {
        float4 segA[kLrfSegBatch + 1];   // per segment: target image x, y, z and (first item - flat position of its first candidate)
        float4 thr[2];                   // screening thresholds by source kind
        LrfSeg seg[kLrfSegBatch];        // item ranges [lo,hi) of the cell rows (+ target image) of this warp's target, as built
        int segI[kLrfSegBatch + 1];      // image code of the segment's row, in place for the queue entry
        int cumE[kLrfSegBatch + 32];     // flat position one past the segment's last candidate
        int queue[64];
        double red[kLrfRaw];
    };
    __shared__ WarpSm wsm[kRowWarps];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int wbase = wid * (int)sizeof(WarpSm);
    // opaque to the compiler: it otherwise rebuilds both from %tid wherever they are used (17 S2R sites in the r04f build)
    asm volatile("" : "+r"(lane), "+r"(wbase));
    WarpSm &W = *reinterpret_cast<WarpSm *>(reinterpret_cast<char *>(wsm) + wbase);
    const int titem = blockIdx.x * kRowWarps + wid;
    if (titem >= D.nunit) return;   // warp-uniform; no block-wide barrier below
    const int t = cell_items[titem];
    if (D.u_excl[t]) return;
    if (GENERAL && D.sharded && D.shard_rows && !row_in_any_shard(D, t)) return;   // another rank's target: stays zero, summed later
    const int ns = D.ncgp_solute;
    const int gt = D.u_grp[t];
    double *lt = lrf + (size_t)QNB_LRF_STRIDE * gt;
    const double cx_ = lt[0], cy_ = lt[1], cz_ = lt[2];
    const double pt[3] = {upos[3 * t], upos[3 * t + 1], upos[3 * t + 2]};
    // any-atom mode: the deciding atoms sit up to rmax2/2 from either switch atom
    const float rl = sqrtf(fmaxf(f32_once(C.rclrf2), 0.f)) + (D.any_atom ? f32_once(C.rmax2) : 0.f);
    const float hi_band = rl * rl * (1.0f + 1e-3f) + 0.05f;   // surely outside the LRF shell above this
    const bool any_all = C.lrf_all[0] || C.lrf_all[1] || C.lrf_all[2] || (D.any_atom && D.use_PBC && C.rclrf2 == -1.0);
    // phi0, phi1, phi2 in FP64 (FP32 phi2 was measured to move E%LRF by 1.1e-6 relative).  phi3 enters only the field,
    // as 1/2 dr.phi3.dr with |dr| ~ 1 A (atom to group centre) against r >= Rc: a (dr/r)^2 ~ 1e-2 correction to
    // phi1, so it is formed and summed in FP32 (relative error ~1e-6 of itself).
    double m[11];
    float h[13];
#pragma unroll
    for (int k = 0; k < 11; k++) m[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 13; k++) h[k] = 0.f;

    // target centre in the frame the plane records are stored in (k_pack_lrf_planes): moved by the image that takes the
    // target's own binned position into the box
    double cw[3] = {cx_, cy_, cz_};
    if (ROWSHIFT)
        for (int d = 0; d < 3; d++) cw[d] -= D.box[d] * floor(pt[d] * D.inv_box[d]);
    // e = item index | image code of the scanned cell row << 26 (ROWSHIFT).  The plane records of a batch are LOADED when
    // the queue fills and USED at the next flush (r04e: 19 % of the stall samples sat on the first use of these loads),
    // the screening steps in between cover their latency.
    double P[12];
    float PQ[3];
    int pend_e = 0, pend_n = 0;
    bool pend = false;   // warp-uniform
    auto plane_load = [&](int e) {
        const int idx = e & 0x3ffffff;
        const char *blk = planes + (size_t)(idx >> 5) * kLrfPlaneBytes;
        const double *pd = reinterpret_cast<const double *>(blk) + (idx & 31);
        const float *pf = reinterpret_cast<const float *>(blk + kLrfPlaneQf) + (idx & 31);
#pragma unroll
        for (int k = 0; k < 12; k++) P[k] = pd[32 * k];   // a run of items reads consecutive words of every plane
#pragma unroll
        for (int k = 0; k < 3; k++) PQ[k] = pf[32 * k];
        pend_n = reinterpret_cast<const int *>(blk + kLrfPlaneCnt)[idx & 31];
        pend_e = e;
    };
    // lrf_update(group1 = source unit of the item, group2 = target): dr = x(i) - cgp_cent(target) - shift
    auto plane_compute = [&]() {
        double ox = cw[0], oy = cw[1], oz = cw[2];
        if (ROWSHIFT) {
            // shift = boxlength*nint((x(cgp(group1)%iswitch) - lrf(group2)%cgp_cent)*inv_boxl): with both ends taken into
            // the box it is minus the image the row was scanned at (|delta| < box/2 for every scanned cell, see qnb.cu)
            const int code = (int)((unsigned)pend_e >> 26);
            ox -= (double)((code & 3) - 1) * D.box[0];
            oy -= (double)(((code >> 2) & 3) - 1) * D.box[1];
            oz -= (double)(((code >> 4) & 3) - 1) * D.box[2];
        }
        lrf_atom_d(m, h, P[0] - ox, P[1] - oy, P[2] - oz, P[3], PQ[0]);
        lrf_atom_d(m, h, P[4] - ox, P[5] - oy, P[6] - oz, P[7], PQ[1]);
        lrf_atom_d(m, h, P[8] - ox, P[9] - oy, P[10] - oz, P[11], PQ[2]);
        if (pend_n > 3) {
            // the rest of a larger solute group from the per-atom records (stored unmoved)
            const int idx = pend_e & 0x3ffffff;
            if (ROWSHIFT) {
                const double4 ip = item_pos[idx];
                ox += D.box[0] * floor(ip.x * D.inv_box[0]); oy += D.box[1] * floor(ip.y * D.inv_box[1]); oz += D.box[2] * floor(ip.z * D.inv_box[2]);
            }
            const int a0 = src_off[idx];
            for (int k = a0 + 3; k < a0 + pend_n; k++) {
                const double4 sa = src[k];
                lrf_atom_d(m, h, sa.x - ox, sa.y - oy, sa.z - oz, sa.w, (float)sa.w);
            }
        }
    };
    auto accumulate = [&](int e) {
        if (ROWSHIFT || !D.use_PBC) { plane_load(e); plane_compute(); return; }
        // small periodic boxes: the shift from the pair itself; the group's switch atom is the unit's binned position
        // (checked at init: g_switch[u_grp[u]] == u_sw[u])
        const int idx = e & 0x3ffffff;
        const double4 ip = item_pos[idx];
        const double ox = cx_ + pshift(ip.x - cx_, D.box[0], D.inv_box[0]);
        const double oy = cy_ + pshift(ip.y - cy_, D.box[1], D.inv_box[1]);
        const double oz = cz_ + pshift(ip.z - cz_, D.box[2], D.inv_box[2]);
        const float oxf = (float)ox, oyf = (float)oy, ozf = (float)oz;
        const int a0 = src_off[idx], a1 = src_off[idx + 1];
        for (int k = a0; k < a1; k++) {
            const double4 sa = src[k];
            const float4 sf = srcf[k];
            lrf_atom(m, h, sa.x - ox, sa.y - oy, sa.z - oz, sa.w, sf.x - oxf, sf.y - oyf, sf.z - ozf, sf.w);
        }
    };
    // a full batch from the queue: the pending one is summed, this one becomes pending
    auto flush = [&](int e) {
        if (ROWSHIFT || !D.use_PBC) {
            if (pend) plane_compute();
            plane_load(e);
            pend = true;
        } else accumulate(e);
    };

    const int cu = cell_of[t];
    const int cx = cu % G.n[0], cy = (cu / G.n[0]) % G.n[1], cz = cu / (G.n[0] * G.n[1]);
    const DimRange rz = dim_range(cz, reach.z, G.n[2], G.periodic), ry = dim_range(cy, reach.y, G.n[1], G.periodic);
    const XSeg xs = x_segments(cx, reach.x, G.n[0], G.periodic);
    // every cell is in reach (sphere with RcLRF spanning it, "no LRF cut-off" boxes): scan the item array as one row
    const bool whole = !ROWSHIFT && rz.count == G.n[2] && ry.count == G.n[1] && xs.n == 1 && xs.lo[0] == 0 && xs.hi[0] == G.n[0];
    const int nrow = whole ? 1 : rz.count * ry.count * (ROWSHIFT ? 2 : xs.n);
    // screening thresholds of this target against solute / water sources (FP32, see screen_r2)
    const bool t_sol = t < ns;
    const int cls_s = t_sol ? 0 : 1, cls_w = t_sol ? 1 : 2;     // class with a solute / a water source
    auto lo_of = [](double c2) { return f32_once(c2) * (1.0f - 1e-3f) - 0.05f; };
    auto hi_of = [](double c2) { return f32_once(c2) * (1.0f + 1e-3f) + 0.05f; };
    const float in_lo_s = lo_of(C.rc2_of(cls_s)), in_hi_s = hi_of(C.rc2_of(cls_s));
    const float in_lo_w = lo_of(C.rc2_of(cls_w)), in_hi_w = hi_of(C.rc2_of(cls_w));
    const float out_lo = lo_of(C.rclrf2), out_hi = hi_of(C.rclrf2);
    // screening thresholds by source kind in shared memory (one 16-byte load per step instead of eight registers)
    float4 *thr = W.thr;
    if (lane == 0) {
        const float kInf = __int_as_float(0x7f800000);
        thr[0] = make_float4(in_lo_w, in_hi_w, C.lrf_all_of(cls_w) ? kInf : out_hi, C.lrf_all_of(cls_w) ? kInf : out_lo);
        thr[1] = make_float4(in_lo_s, in_hi_s, C.lrf_all_of(cls_s) ? kInf : out_hi, C.lrf_all_of(cls_s) ? kInf : out_lo);
    }
    __syncwarp();
    const float bx = f32_once(D.box[0]), by = f32_once(D.box[1]), bz = f32_once(D.box[2]);
    const float ibx = f32_once(D.inv_box[0]), iby = f32_once(D.inv_box[1]), ibz = f32_once(D.inv_box[2]);
    // target position as the candidates are stored: wrapped into the box when periodic (k_pack_items)
    double ptw[3] = {pt[0], pt[1], pt[2]};
    if (G.periodic)
        for (int d = 0; d < 3; d++) ptw[d] -= D.box[d] * floor(ptw[d] * D.inv_box[d]);
    const float ptf[3] = {f32_once(ptw[0]), f32_once(ptw[1]), f32_once(ptw[2])};
    LrfSeg *seg = W.seg;
    int *myq = W.queue;
    const unsigned lt_mask = (1u << lane) - 1u;
    int qn = 0;   // entries waiting in this warp's queue (warp-uniform)
    // The candidates of all segments of a batch of rows are walked as ONE sequence of 32-wide steps (r04f: with steps cut at
    // the segment ends 43 % of the lanes were idle, a segment holds ~30 candidates).  The lanes of a step find their
    // segment from the segment ends: lane j looks at the end of segment kmin + j, the ends that fall inside the step are
    // OR-ed into a bit mask and a lane's segment is kmin + the number of ends at or before its position.
    struct StepRegs { float4 pf; int k, ent; };   // screening record, segment, queue entry of a lane's candidate
    int kmin = 0, flat_total = 0;
    auto fetch = [&](int sidx, StepRegs &r) {
        const int f0 = 32 * sidx;
        const int rel = W.cumE[kmin + lane] - f0;
        const unsigned mask = __reduce_or_sync(kFull, (rel >= 1 && rel <= 31) ? (1u << rel) : 0u);
        const int k = kmin + __popc(mask & (0xffffffffu >> (31 - lane)));
        kmin += __popc(mask) + (__any_sync(kFull, rel == 32) ? 1 : 0);
        const int idx = f0 + lane + __float_as_int(W.segA[k].w);
        r.k = k;
        r.ent = -1;
        r.pf = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));   // lanes past the last candidate: the "no source" id
        if (f0 + lane < flat_total) { r.ent = idx | W.segI[k]; r.pf = item_posf[idx]; }
    };
    // one 32-wide step: FP32 screening of the candidates against the target image of their rows, the FP64 test for the
    // few inside the safety band, accepted items to the queue (or straight to the accumulation)
    auto process = [&](const StepRegs &r) {
        const float4 pf = r.pf;
        const float4 st = W.segA[r.k];
        const int ent = r.ent;
        const int idx = ent & 0x3ffffff;
        const int s = __float_as_int(pf.w);
        float dx = pf.x - st.x, dy = pf.y - st.y, dz = pf.z - st.z;
        if (!ROWSHIFT && D.use_PBC) { dx -= bx * rintf(dx * ibx); dy -= by * rintf(dy * iby); dz -= bz * rintf(dz * ibz); }
        const float r2f = dx * dx + dy * dy + dz * dz;
        const bool s_sol = s < ns;
        // "no LRF cut-off" classes carry infinite outer thresholds
        const float4 th = thr[s_sol ? 1 : 0];     // {in_lo, in_hi, out_hi, out_lo} of the source's kind
        bool rej = r2f < th.x || r2f > th.z;      // listed, or outside the shell
        bool sure = r2f > th.y && r2f < th.w;     // surely inside the shell; neither: the FP64 test decides
        if (GENERAL && D.any_atom && (t_sol || s_sol)) { rej = !any_all && r2f > hi_band; sure = false; }
        rej = rej || s < 0 || s == t;
        bool accept = sure && !rej;
        const bool slow = !rej && ((GENERAL && D.sharded) || !sure);
        if (__any_sync(kFull, slow)) {
            if (slow) {
                bool owner_is_t;
                const int cls = pair_class(t, s, ns, owner_is_t);
                if (GENERAL && D.sharded && !(D.shard_rows ? row_in_shard(D, cls, t) : in_shard(D, cls, owner_is_t ? t : s))) accept = false;
                else if (!sure) {
                    const double4 ip = item_pos[idx];
                    const double ps[3] = {ip.x, ip.y, ip.z};
                    // outside the class cut-off (else: a listed pair) and inside the LRF cut-off
                    accept = (owner_is_t ? unit_pair_test(D, C, x, cls, t, s, pt, ps) : unit_pair_test(D, C, x, cls, s, t, ps, pt)).lrf;
                }
            }
        }
        if (!COMPACT) {
            // nearly every scanned unit is a source (sphere with RcLRF covering it): no queueing needed
            if (accept) accumulate(ent);
        } else {
            const unsigned mask = __ballot_sync(kFull, accept);
            if (accept) myq[qn + __popc(mask & lt_mask)] = ent;
            qn += __popc(mask);
            if (qn >= 32) {
                __syncwarp();
                flush(myq[lane]);
                const int rest = qn - 32;
                int moved = 0;
                if (lane < rest) moved = myq[32 + lane];
                __syncwarp();
                if (lane < rest) myq[lane] = moved;
                qn = rest;
            }
        }
    };
    for (int r0 = 0; r0 < nrow; r0 += kLrfSegBatch) {
        __syncwarp();
        int nseg = 0;
        if (whole) { if (lane == 0) seg[0] = LrfSeg{0, D.nunit, ptf[0], ptf[1], ptf[2], 0}; nseg = 1; }
        else for (int rb = r0; rb < min(nrow, r0 + kLrfSegBatch); rb += 32) {
            const int r = rb + lane;
            LrfSeg sg{0, 0, ptf[0], ptf[1], ptf[2], 0};
            if (r < min(nrow, r0 + kLrfSegBatch)) {
            const int nxs = ROWSHIFT ? 2 : xs.n;
            const int sgi = r % nxs, iy = (r / nxs) % ry.count, iz = r / (nxs * ry.count);
            const int zu = rz.start + iz, yu = ry.start + iy;           // unwrapped cell indices of the row
            const int z = (zu % G.n[2] + G.n[2]) % G.n[2], y = (yu % G.n[1] + G.n[1]) % G.n[1];
            const int rowbase = (z * G.n[1] + y) * G.n[0];
            if (!ROWSHIFT) {
                sg.lo = cell_start[rowbase + (sgi == 0 ? xs.lo[0] : xs.lo[1])];
                sg.hi = cell_start[rowbase + (sgi == 0 ? xs.hi[0] : xs.hi[1])];
            } else {
                // distance from the target to the row's y/z slab (cell size = box/n exactly, see cell_coord)
                const float cyl = by / (float)G.n[1], czl = bz / (float)G.n[2], cxl = bx / (float)G.n[0];
                const float ylo = (float)yu * cyl, zlo = (float)zu * czl;
                const float dy = fmaxf(0.f, fmaxf(ylo - ptf[1], ptf[1] - (ylo + cyl)));
                const float dz = fmaxf(0.f, fmaxf(zlo - ptf[2], ptf[2] - (zlo + czl)));
                const float rr = rl * (1.0f + 1e-3f) + 0.05f;             // same safety as hi_band
                const float h2 = rr * rr - dy * dy - dz * dz;
                if (h2 > 0.f) {
                    const float xh = sqrtf(h2) + 1e-3f * cxl;
                    int xa = max(cx - reach.x, (int)floorf((ptf[0] - xh) / cxl));
                    int xb = min(cx + reach.x, (int)floorf((ptf[0] + xh) / cxl));
                    // unwrapped range [xa,xb] -> part in the image of xa (sgi 0) and the rest (sgi 1)
                    const int ia = (xa >= 0 ? xa / G.n[0] : -((-xa + G.n[0] - 1) / G.n[0]));   // floor(xa/n)
                    const int split = (ia + 1) * G.n[0];                                          // first cell of the next image
                    int a0 = sgi == 0 ? xa : max(xa, split), b0 = sgi == 0 ? min(xb, split - 1) : xb;
                    if (a0 <= b0) {
                        const int img = sgi == 0 ? ia : ia + 1;
                        sg.lo = cell_start[rowbase + (a0 - img * G.n[0])];
                        sg.hi = cell_start[rowbase + (b0 - img * G.n[0]) + 1];
                        const int iyg = (yu >= 0 ? yu / G.n[1] : -((-yu + G.n[1] - 1) / G.n[1]));
                        const int izg = (zu >= 0 ? zu / G.n[2] : -((-zu + G.n[2] - 1) / G.n[2]));
                        // candidate image = stored position + img*box: subtract it from the target instead
                        sg.tx = ptf[0] - (float)img * bx; sg.ty = ptf[1] - (float)iyg * by; sg.tz = ptf[2] - (float)izg * bz;
                        sg.img = (img + 1) | ((iyg + 1) << 2) | ((izg + 1) << 4);
                    }
                }
            }
            }
            // only rows that hold candidates, appended in lane order
            const unsigned keep = __ballot_sync(kFull, sg.hi > sg.lo);
            if (sg.hi > sg.lo) seg[nseg + __popc(keep & ((1u << lane) - 1u))] = sg;
            nseg += __popc(keep);
        }
        __syncwarp();
        // flat positions of the segments' candidates (exclusive scan of the lengths)
        int carry = 0;
        for (int g0 = 0; g0 < nseg; g0 += 32) {
            const int k = g0 + lane;
            LrfSeg ms{0, 0, 0.f, 0.f, 0.f, 0};
            if (k < nseg) ms = seg[k];
            const int len = ms.hi - ms.lo;
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            const int end = carry + incl;
            if (k < nseg) {
                W.segA[k] = make_float4(ms.tx, ms.ty, ms.tz, __int_as_float(ms.lo - (end - len)));
                W.segI[k] = ROWSHIFT ? (ms.img << 26) : 0;
            }
            W.cumE[k] = k < nseg ? end : 0x3fffffff;
            carry = __shfl_sync(kFull, end, 31);
        }
        for (int k = ((nseg + 31) & ~31) + lane; k < nseg + 32; k += 32) W.cumE[k] = 0x3fffffff;
        if (lane == 0) { W.segA[nseg] = make_float4(0.f, 0.f, 0.f, 0.f); W.segI[nseg] = 0; }
        __syncwarp();
        flat_total = carry;
        kmin = 0;
        const int nsteps = (flat_total + 31) >> 5;
        if (nsteps > 0) {
            // the records of the next two steps are in flight while the current one is tested (r04g: with one step of lead
            // and full steps the record gather was the largest stall); QNB_LRF_DEPTH register sets rotated by copies, one instance
            // of the step body (four unrolled instances ran into instruction-cache misses, r04c)
            StepRegs r[QNB_LRF_DEPTH];
            fetch(0, r[0]);
#pragma unroll
            for (int d = 1; d < QNB_LRF_DEPTH - 1; d++) {
                r[d] = r[d - 1];
                if (d < nsteps) fetch(d, r[d]);
            }
            for (int sidx = 0; sidx < nsteps; sidx++) {
                r[QNB_LRF_DEPTH - 1] = r[QNB_LRF_DEPTH - 2];
                if (sidx + QNB_LRF_DEPTH - 1 < nsteps) fetch(sidx + QNB_LRF_DEPTH - 1, r[QNB_LRF_DEPTH - 1]);
                process(r[0]);
#pragma unroll
                for (int d = 0; d < QNB_LRF_DEPTH - 1; d++) r[d] = r[d + 1];
            }
        }
    }
    __syncwarp();
    if (pend) plane_compute();
    if (lane < qn) accumulate(myq[lane]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 11; k++) {
        const double a = warp_sum(m[k]);
        if (lane == 0) W.red[k] = a;
    }
#pragma unroll
    for (int k = 0; k < 13; k++) {
        const double a = warp_sum((double)h[k]);
        if (lane == 0) W.red[11 + k] = a;
    }
    __syncwarp();
    // LRF_TYPE order after cgp_cent: phi0, phi1(3), phi2(a)%b (9), phi3(3*(n-1)+j)%k (27) from the unique moments
    for (int k = lane; k < 40; k += 32) lt[3 + k] = lrf_finish(W.red, kLrfExpand[k]);
}

// ---------------------------------------------------------------- LRF, sphere with the LRF cut-off spanning it
// Every unit is a source of (nearly) every target: the classic all-pairs tiling.  One THREAD per target keeps its 20
// moments in registers; a block streams a slice of the cell-ordered source items through shared memory (positions,
// atom ranges, atoms), so every lane reads the same source atom (a broadcast, no gathers, no dependent global loads)
// and the moments of the slice are added to an unexpanded [nunit][20] buffer that k_lrf_expand turns into LRF_TYPE.
// Switching-atom lists, unsharded, non-periodic only (the general kernel above serves everything else).
constexpr int kLrfTileItems = 64, kLrfTileAtoms = 320;
template <int MINB>   // resident blocks per SM the register allocation aims at (5: 96 registers, 6: 80)
__global__ void __launch_bounds__(128, MINB)
k_lrf_allpairs(Dev D, Cut C, const double *__restrict__ x, const double *__restrict__ upos, const double4 *__restrict__ item_pos,
               const float4 *__restrict__ item_posf, const int *__restrict__ src_off, const double4 *__restrict__ src,
               const double *__restrict__ lrf, double *__restrict__ mom /* [nunit][kLrfRaw] raw sums */) {
    __shared__ float4 s_pos[kLrfTileItems];
    __shared__ int s_a0[kLrfTileItems + 1];
    __shared__ double4 s_atom[kLrfTileAtoms];
    __shared__ float4 s_atomf[kLrfTileAtoms];   // FP32 copy for the phi3 path and the rsqrt seed: conversions cost XU slots
    __shared__ int s_tile_end;
    const int ns = D.ncgp_solute;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < D.nunit && !D.u_excl[t];
    const int tt = live ? t : 0;
    const int gt = D.u_grp[tt];
    const double cx_ = lrf[(size_t)QNB_LRF_STRIDE * gt], cy_ = lrf[(size_t)QNB_LRF_STRIDE * gt + 1], cz_ = lrf[(size_t)QNB_LRF_STRIDE * gt + 2];
    const double pt[3] = {upos[3 * tt], upos[3 * tt + 1], upos[3 * tt + 2]};
    const float ptf[3] = {f32_once(pt[0]), f32_once(pt[1]), f32_once(pt[2])};
    const float cxf = f32_once(cx_), cyf = f32_once(cy_), czf = f32_once(cz_);
    const bool t_sol = tt < ns;
    const int cls_s = t_sol ? 0 : 1, cls_w = t_sol ? 1 : 2;
    auto lo_of = [](double c2) { return f32_once(c2) * (1.0f - 1e-3f) - 0.05f; };
    auto hi_of = [](double c2) { return f32_once(c2) * (1.0f + 1e-3f) + 0.05f; };
    const float in_lo_s = lo_of(C.rc2_of(cls_s)), in_hi_s = hi_of(C.rc2_of(cls_s));
    const float in_lo_w = lo_of(C.rc2_of(cls_w)), in_hi_w = hi_of(C.rc2_of(cls_w));
    const float out_lo = lo_of(C.rclrf2), out_hi = hi_of(C.rclrf2);
    double m[11];
    float h[13];
#pragma unroll
    for (int k = 0; k < 11; k++) m[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 13; k++) h[k] = 0.f;
    // this block's slice of the items
    const int per = (D.nunit + gridDim.y - 1) / gridDim.y;
    const int lo = blockIdx.y * per, hi = min(D.nunit, lo + per);
    for (int base = lo; base < hi;) {
        __syncthreads();
        // tile = up to kLrfTileItems items whose atoms fit
        if (threadIdx.x == 0) {
            int e = min(hi, base + kLrfTileItems);
            const int a_first = src_off[base];
            while (e > base + 1 && src_off[e] - a_first > kLrfTileAtoms) e--;
            s_tile_end = e;
        }
        __syncthreads();
        const int tend = s_tile_end, nit = tend - base;
        const int a_first = src_off[base], nat = src_off[tend] - a_first;
        for (int k = threadIdx.x; k < nit; k += blockDim.x) s_pos[k] = item_posf[base + k];
        for (int k = threadIdx.x; k <= nit; k += blockDim.x) s_a0[k] = src_off[base + k] - a_first;
        for (int k = threadIdx.x; k < min(nat, kLrfTileAtoms); k += blockDim.x) {
            const double4 a = src[a_first + k];
            s_atom[k] = a;
            s_atomf[k] = make_float4((float)a.x, (float)a.y, (float)a.z, (float)a.w);
        }
        __syncthreads();
        if (live && nat <= kLrfTileAtoms) {
            for (int it = 0; it < nit; it++) {
                const float4 pf = s_pos[it];
                const int s = __float_as_int(pf.w);
                if (s < 0 || s == t) continue;
                const float dx = pf.x - ptf[0], dy = pf.y - ptf[1], dz = pf.z - ptf[2];
                const float r2f = dx * dx + dy * dy + dz * dz;
                const bool s_sol = s < ns;
                const float il = s_sol ? in_lo_s : in_lo_w, ih = s_sol ? in_hi_s : in_hi_w;
                // zones: 0 listed or outside the shell, 1 surely inside it, 2 the FP64 test decides
                int zone = (r2f < il || r2f > out_hi) ? 0 : (r2f > ih && r2f < out_lo) ? 1 : 2;
                if (zone == 2) {
                    bool owner_is_t;
                    const int cls = pair_class(t, s, ns, owner_is_t);
                    const double4 ip = item_pos[base + it];
                    const double ps[3] = {ip.x, ip.y, ip.z};
                    const bool lrf_pair = (owner_is_t ? unit_pair_test(D, C, x, cls, t, s, pt, ps)
                                                      : unit_pair_test(D, C, x, cls, s, t, ps, pt)).lrf;
                    zone = lrf_pair ? 1 : 0;
                }
                if (zone != 1) continue;
                // lrf_update (nonbondene.f90:656-719), see lrf_atom; FP32 displacement from FP32 copies
                // (|x| ~ 30 A: 2e-6 A, the phi3 path is FP32 anyway)
                const int k0 = s_a0[it], k1 = s_a0[it + 1];   // the same for every lane
                if (k1 - k0 == 3) {
                    // three-atom unit (every water, most solute groups): straight-line code, the three dependency chains
                    // interleave (one atom at a time left ~2 cycles of fixed-latency stall per instruction, r03h)
                    const double4 s0 = s_atom[k0], s1 = s_atom[k0 + 1], s2 = s_atom[k0 + 2];
                    const float4 f0 = s_atomf[k0], f1 = s_atomf[k0 + 1], f2 = s_atomf[k0 + 2];
                    lrf_atom(m, h, s0.x - cx_, s0.y - cy_, s0.z - cz_, s0.w, f0.x - cxf, f0.y - cyf, f0.z - czf, f0.w);
                    lrf_atom(m, h, s1.x - cx_, s1.y - cy_, s1.z - cz_, s1.w, f1.x - cxf, f1.y - cyf, f1.z - czf, f1.w);
                    lrf_atom(m, h, s2.x - cx_, s2.y - cy_, s2.z - cz_, s2.w, f2.x - cxf, f2.y - cyf, f2.z - czf, f2.w);
                } else
                    for (int k = k0; k < k1; k++) {
                        const double4 sa = s_atom[k];
                        const float4 af = s_atomf[k];
                        lrf_atom(m, h, sa.x - cx_, sa.y - cy_, sa.z - cz_, sa.w, af.x - cxf, af.y - cyf, af.z - czf, af.w);
                    }
            }
        }
        base = tend;
    }
    if (live) {
        double *mt = mom + (size_t)kLrfRaw * t;
#pragma unroll
        for (int k = 0; k < 11; k++) atomicAdd(&mt[k], m[k]);
#pragma unroll
        for (int k = 0; k < 13; k++) atomicAdd(&mt[11 + k], (double)h[k]);
    }
}
// raw sums -> LRF_TYPE (phi0, phi1(3), phi2(3x3), phi3(9x3)) of the unit's charge group
__global__ void k_lrf_expand(Dev D, const double *__restrict__ mom, double *__restrict__ lrf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.nunit * 40) return;
    const int t = i / 40, k = i - 40 * t;
    if (D.u_excl[t]) return;
    lrf[(size_t)QNB_LRF_STRIDE * D.u_grp[t] + 3 + k] = lrf_finish(mom + (size_t)kLrfRaw * t, kLrfExpand[k]);
}

}  // namespace qnb
