// qnb.cu -- C ABI (include/qnb.h) of the B200 nonbonded engine: device state, list build and
// step orchestration.  Kernels live in qnb_lists.cuh / qnb_forces.cuh, host tables in qnb_tables.hpp.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qnb.h"
#include "qnb_kernels.cuh"
#include "qnb_forces.cuh"
#include "qnb_rows.cuh"
#include "qnb_shake.cuh"
#include "qnb_lists.cuh"
#include "qnb_tables.hpp"

namespace qnb {

static thread_local std::string g_err;
static int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- minimal NCCL binding, resolved at run time from whichever libnccl the process already has
struct ncclUniqueIdT { char internal[128]; };
typedef void *ncclCommT;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueIdT *) = nullptr;
    int (*CommInitRank)(ncclCommT *, int, ncclUniqueIdT, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclCommT, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclCommT) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        GetUniqueId = (int (*)(ncclUniqueIdT *))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(ncclCommT *, int, ncclUniqueIdT, int))dlsym(lib, "ncclCommInitRank");
        AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclCommT, cudaStream_t))dlsym(lib, "ncclAllReduce");
        CommDestroy = (int (*)(ncclCommT))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
    }
};
static NcclApi g_nccl;
constexpr int kNcclDouble = 8, kNcclSum = 0;

template <typename T>
struct DBuf {   // device buffer that only grows
    T *p = nullptr;
    size_t cap = 0;
    bool view = false;   // points into another allocation (the arena shared with peer GPUs): never freed or regrown here
    void alias(T *q, size_t n) { p = q; cap = n; view = true; }
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (view) return fail("internal: a view buffer cannot grow (%zu > %zu)", n, cap);
        if (p) cudaFree(p);
        p = nullptr;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { cap = 0; return fail("cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; cap = 0; view = false; }
};

template <typename T>
static int upload(DBuf<T> &b, const std::vector<T> &v) {
    if (b.ensure(std::max<size_t>(v.size(), 1))) return 1;
    if (!v.empty()) CU(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace qnb

using namespace qnb;

constexpr int kAux = 6;   // auxiliary streams: the kernels of one evaluation run concurrently
namespace qnb { struct BatchSlot; }

struct qnb_handle {
    int device = 0, nsm = 148;
    HostTables T;
    Dev D{};
    cudaStream_t st = nullptr, aux[kAux] = {}, lrf_st = nullptr;
    cudaEvent_t ev_lrf = nullptr;   // end of the LRF accumulation of a list build (lrf_st)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_pack = nullptr, ev_join[kAux] = {};
    cudaEvent_t ev_bt[6] = {};   // qnb_bench_build_lists: LRF kernels (0,1), row count pass (2,3), row fill pass (4,5)
    bool time_build = false;
    float build_ms[3] = {0, 0, 0};   // of the last timed build: LRF accumulation, row count pass, row fill pass
    cudaGraphExec_t graph[2][16] = {};   // [with copies][graph_index(flags)]
    bool use_graph = true, multi_stream = true;
    int water_blocks = 0, solute_blocks = 0;   // blocks per SM of the two persistent kernels (0: what the occupancy allows)
    int graph_launches[2][16] = {};
    bool graph_dirty[2][16] = {};   // set by drop_graphs at init
    // static device tables
    DBuf<double> crg, ljd;
    DBuf<float> crgf, ljf;
    DBuf<int> ctype, grp_of_atom, g_first, g_n, g_switch, g_atoms, g_nq, u_sw, u_grp, sp_off, sp_partner, gs_off,
        gs_atoms, iqseq, nq_off, nq_atoms;
    DBuf<uint8_t> is_q, excl, qbonded, u_excl, ljcode, sp_code;
    DBuf<QPar4> qp_tab, qw_tab;
    DBuf<float4> qp_tabf, qw_tabf;
    DBuf<QStatic> qstatic;
    int n_qq = 0, n_qstatic = 0;
    // per-step
    DBuf<double> x /* x[3n] | lambda[kMaxStates] */, out /* grad[3n] | E[7] | EQ[6*nstates] */, lrf;
    double *hx = nullptr, *hout = nullptr, *hlam = nullptr;   // pinned staging; hlam = tail of hx: [x | lambda] goes up in one copy
    double *lam_dev = nullptr;                                // device lambda = tail of the x buffer
    size_t nout = 0;
    int nE = 0;   // energies per slot: E[7] | EQ[6*nstates]; the output buffer holds kESlots partial copies
    // per-build
    Cut cut{};
    Grid grid{};
    bool have_grid = false;
    DBuf<double> upos;
    DBuf<int> cell_of, cell_count, cell_start, cell_items, counts, row_tot, row_off, flag, pos, flag2, pos2, qp_list, qw_list,
        qp_shift_atom;
    DBuf<uint32_t> rows;
    DBuf<double4> item_pos, src;
    DBuf<char> lrf_planes;   // the LRF sources in planes per 32 items (k_pack_lrf_planes)
    DBuf<float4> item_posf, item_scr, srcf;
    DBuf<int> cell_unsorted, item_nq, src_off, pk_atom, pk_ct;
    DBuf<float> pk_q;
    DBuf<double> pk_qd, px, py, pz;
    int npk = 0;   // packed atoms: non-Q atoms of non-excluded units in cell order
    DBuf<int> nch, choff, ucost, cost_off, wstart_w, wstart_s;   // chunk counts / offsets / costs per unit, warp shares
    int wgrid = 0, sgrid = 0;   // grids of the persistent force kernels (one resident wave)
    int occ_w = 0, occ_s = 0;
    DBuf<int2> wdesc, sdesc;
    DBuf<uint32_t> wrow, srow;
    DBuf<uint16_t> sspec;
    // solvent restraints (restrain_solvent / watpol)
    qnb_solvent_restraints rst{};
    bool rst_set = false;
    double theta_corr[QNB_MAX_SHELLS] = {};
    DBuf<double> wp_theta, wp_shell_theta;
    DBuf<double> lrf_mom;   // [nunit][20] unexpanded LRF moments (all-pairs kernel)
    DBuf<int> wp_shell_n, wp_shell_list;
    double last_rst[kRstOut] = {};
    // MC_volume: state of the last list build and its saved copy
    DBuf<double> x_saved, lrf_saved;
    std::vector<double> hx_build, hx_saved;
    double box_saved[3] = {0, 0, 0}, inv_saved[3] = {0, 0, 0};
    Cut cut_build{}, cut_saved{};
    bool have_saved = false, restoring = false;   // solute chunks: per entry, special-pair codes of the tile atoms
    int nwchunk = 0, nschunk = 0;   // water-row / solute-row chunks
    // round-2 row kernels (qnb_rows.cuh): per-step packed records, pair tables, flat energy lists
    DBuf<int> upk, pk_sw, e_cnt, e_off;
    DBuf<int4> rec_i;
    DBuf<float4> rec_f, wT, wown, pw12;
    DBuf<float2> ljp, pw0;
    DBuf<int2> ww_pairs, pp_pairs, pw_pairs;
    DBuf<double2> wd;
    int n_ww_e = 0, n_pp_e = 0, n_pw_e = 0;
    int occ_wr = 0, occ_sr = 0, wrows_minb = 5;
    bool legacy_rows = false, pw_hlj = false;
    RowPar rowpar{};
    EnergyPar epar{};
    FixFrame fix{};
    int nqp = 0, nqw = 0;
    bool qp_done = false, qw_done = false, lists_built = false;
    bool x_from_nonbond = false;   // the device coordinates are the ones of the last qnb_nonbond (qnb_shake with xx == NULL)
    DBuf<int> bead_atoms;          // qnb_qcp_beads work buffers, kept between calls
    DBuf<double> bead_base, bead_disp, bead_eq;
    int64_t total_rows = 0;
    int3 lrf_reach{1, 1, 1}, list_reach{1, 1, 1};
    double box[3] = {0, 0, 0}, inv_box[3] = {0, 0, 0};
    // SHAKE of solvent-sized molecules (qnb_set_constraints / qnb_shake)
    DBuf<int> shk_first;
    DBuf<ShakePair> shk_ij;
    DBuf<double> shk_d2, shk_winv, shk_x, shk_xx;
    DBuf<unsigned long long> shk_iter;   // [0] summed sweeps, [1] (as int) failure flag
    int shk_nmol = 0;
    bool shk_set = false;
    // comm
    ncclCommT comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory all-reduce (qnb_comm_ipc_export / qnb_comm_ipc_attach): [out | lrf | control words] live in ONE
    // allocation that the other ranks of the node map through CUDA IPC
    double *arena = nullptr;
    size_t arena_doubles = 0, arena_lrf_off = 0, arena_ctl_off = 0;
    P2PComm p2p{};
    void *peer_base[kMaxPeers] = {};
    // caller buffers used directly by the step's copies (page-locked on first sight, cudaHostRegister): no staging copy of
    // x, and the gradient is added into the caller's d by the device (k_add_out over mapped host memory)
    struct HostReg { const void *p; size_t bytes; void *dev; };
    std::vector<HostReg> hostregs;
    bool host_direct = true;
    const double *x_direct = nullptr;      // this call's registered x (nullptr: staged through hx)
    double *d_direct = nullptr;            // device alias of this call's registered d (nullptr: added on the host)
    const double *graph_x[16] = {};        // what the with-copies graphs were captured with
    double *graph_d[16] = {};
    bool d_overwrite = false, graph_dow[16] = {};   // QNB_FLAG_D_IS_ZERO: copy into the registered d instead of adding
    double *d_host = nullptr;              // the caller's registered d itself (host address), target of that copy
    int share = 1;                         // systems that advance together on this GPU (qnb_build_lists_batch): grids are sized
                                           // for 1/share of the SMs so that the kernels of the systems run side by side
    qnb::BatchSlot *collect = nullptr;     // set while a batch plan records this handle's launch arguments
    int64_t md_phase = 0;                  // qnb_bench_md: MD step counter, runs on across calls
    uint64_t build_serial = 0;             // counts list builds and box changes: batch plans bake row pointers and sizes
    // stats
    int64_t launches = 0, last_h2d = 0, last_d2h = 0;
    double t_stage_in = 0, t_issue = 0, t_wait = 0, t_add_out = 0;   // host-side seconds of the last qnb_nonbond
    DBuf<char> flush;
    int last_flags = 0;
};

namespace qnb {

// ---- one launch for several systems (k_batched, qnb_kernels.cuh): the launch sites of launch_step_kernel either launch
// their kernel for one handle or, when a collector is installed, record the argument tuple of this handle; the batch
// plan then launches every kernel type once for all of its handles.
struct BatchSlot {
    std::vector<unsigned char> args;   // W argument tuples, raw bytes
    size_t stride = 0, smem = 0;
    std::vector<int2> grids;
    dim3 grid{0, 0, 1};
    void (*launch)(const void *, const int2 *, dim3, size_t, cudaStream_t) = nullptr;
    bool bad = false;                  // the handles do not agree on the kernel instantiation
    DBuf<unsigned char> dargs;
    DBuf<int2> dgrids;
    void reset() { args.clear(); grids.clear(); grid = dim3(0, 0, 1); smem = 0; launch = nullptr; bad = false; stride = 0; }
};
template <class Body, int THREADS, int MINB, class... P>
static void launch_batched(const void *dargs, const int2 *dgrids, dim3 grid, size_t smem, cudaStream_t st) {
    k_batched<Body, THREADS, MINB, P...><<<grid, THREADS, smem, st>>>(static_cast<const cuda::std::tuple<P...> *>(dargs), dgrids);
}
template <class Body, int THREADS, int MINB, class... P>
static void collect_launch(BatchSlot &s, dim3 grid, size_t smem, P... a) {
    using T = cuda::std::tuple<P...>;
    static_assert(sizeof(T) % 4 == 0, "argument tuple");
    auto fn = &launch_batched<Body, THREADS, MINB, P...>;
    if (s.launch && s.launch != fn) { s.bad = true; return; }
    s.launch = fn;
    s.stride = sizeof(T);
    T t(a...);
    const unsigned char *b = reinterpret_cast<const unsigned char *>(&t);
    s.args.insert(s.args.end(), b, b + sizeof(T));
    s.grids.push_back(make_int2((int)grid.x, (int)grid.y));
    s.grid.x = std::max(s.grid.x, grid.x); s.grid.y = std::max(s.grid.y, grid.y);
    s.smem = std::max(s.smem, smem);
}

#define LAUNCH_ON(h, stream, kernel, grid, block, smem, ...)            \
    do {                                                               \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
        (h)->launches++;                                               \
    } while (0)
#define LAUNCH(h, kernel, grid, block, smem, ...) LAUNCH_ON(h, (h)->st, kernel, grid, block, smem, __VA_ARGS__)
// a step kernel: launched for this handle, or its arguments recorded for a batched launch (h->collect)
#define STEP_LAUNCH(h, stream, BODY, THREADS, MINB, kernel, grid, smem, ...)                                   \
    do {                                                                                                       \
        if ((h)->collect) collect_launch<BODY, THREADS, MINB>(*(h)->collect, dim3(grid), smem, __VA_ARGS__);   \
        else LAUNCH_ON(h, stream, kernel, grid, THREADS, smem, __VA_ARGS__);                                   \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// fixed-point frame of the row kernels: the box when periodic, else a power-of-two period that covers the binned extent
// four times over (only differences between listed partners are ever formed)
static void refresh_frame(qnb_handle *h) {
    const Grid &G = h->grid;
    FixFrame &F = h->fix;
    for (int d = 0; d < 3; d++) {
        double period;
        if (h->D.use_PBC) { F.org[d] = 0.0; period = h->box[d] > 0 ? h->box[d] : 1.0; }
        else {
            const double ext = G.inv_box[d];   // extent of the padded bounding box (make_grid)
            period = 128.0;
            while (period < 4.0 * (ext + 16.0)) period *= 2.0;
            F.org[d] = G.org[d] - 8.0;
        }
        F.inv_period[d] = 1.0 / period;
        F.scale[d] = (float)(period / 4294967296.0);
        h->rowpar.scale[d] = F.scale[d];
        h->epar.box[d] = h->box[d]; h->epar.inv_box[d] = h->inv_box[d];
    }
}
static void refresh_dev(qnb_handle *h) {
    for (int d = 0; d < 3; d++) { h->D.box[d] = h->box[d]; h->D.inv_box[d] = h->inv_box[d]; }
    if (h->have_grid) refresh_frame(h);
}

static int init_device(qnb_handle *h) {
    HostTables &T = h->T;
    const qnb_system &s = T.s;
    Dev &D = h->D;
    D.natom = s.natom; D.nat_solute = s.nat_solute; D.nwat = s.nwat; D.ncgp = s.ncgp; D.ncgp_solute = s.ncgp_solute;
    D.nunit = T.nunit; D.nqat = s.nqat; D.nstates = s.nstates; D.nct = T.nct;
    D.use_PBC = s.use_PBC != 0; D.use_LRF = s.use_LRF != 0; D.geometric = s.ivdw_rule == QNB_VDW_GEOMETRIC;
    D.spc_water = (s.ivdw_rule == QNB_VDW_GEOMETRIC) && (s.solvent_type == QNB_SOLVENT_SPC);   // potene.f90:347
    D.qswitch0 = s.qswitch - 1;
    D.any_atom = s.iuse_switch_atom != 1;
    D.sharded = !(s.pp_start <= 1 && s.pp_end >= s.ncgp_solute && s.pw_start <= 1 && s.pw_end >= s.ncgp_solute &&
                  s.ww_start <= 1 && s.ww_end >= s.nwat);
    // Sharded builds partition ROWS (row_in_shard; shards the candidate scan as well as the pairs: r02d, 2 x B200, list
    // build 3.54 -> 2.39 ms, step 0.220 -> 0.174 ms against the pair partition).  QNB_SHARD_PAIRS=1 selects the reference's
    // pair partition (every rank holds exactly the pairs of its calculation_assignment ranges).
    D.shard_rows = getenv("QNB_SHARD_PAIRS") ? 0 : 1;
    D.el14 = s.el14_scale; D.el14f = (float)s.el14_scale;
    for (int d = 0; d < 3; d++) D.xpcent[d] = s.xpcent[d];
    D.pp_s = s.pp_start; D.pp_e = s.pp_end; D.pw_s = s.pw_start; D.pw_e = s.pw_end; D.qp_s = s.qp_start; D.qp_e = s.qp_end;
    D.ww_s = s.ww_start; D.ww_e = s.ww_end; D.qw_s = s.qw_start; D.qw_e = s.qw_end; D.at_s = s.natom_start; D.at_e = s.natom_end;

    std::vector<float> crgf(T.crg.begin(), T.crg.end());
    std::vector<float> ljf((size_t)T.nct * 6);
    for (int t = 0; t < T.nct; t++)
        for (int c = 0; c < 3; c++) { ljf[(t * 3 + c) * 2] = (float)T.lj_a[t * 3 + c]; ljf[(t * 3 + c) * 2 + 1] = (float)T.lj_b[t * 3 + c]; }
    std::vector<double> ljd(ljf.begin(), ljf.end());
    for (int t = 0; t < T.nct; t++)
        for (int c = 0; c < 3; c++) { ljd[(t * 3 + c) * 2] = T.lj_a[t * 3 + c]; ljd[(t * 3 + c) * 2 + 1] = T.lj_b[t * 3 + c]; }
    std::vector<int> g_nq(s.ncgp, 0);
    for (int g = 0; g < s.ncgp; g++)
        for (int k = 0; k < T.g_n[g]; k++) g_nq[g] += !T.is_q[T.g_atoms[T.g_first[g] + k]];
    if (upload(h->crg, T.crg) || upload(h->crgf, crgf) || upload(h->ljf, ljf) || upload(h->ljd, ljd) || upload(h->ctype, T.ctype) ||
        upload(h->grp_of_atom, T.grp_of_atom) || upload(h->g_first, T.g_first) || upload(h->g_n, T.g_n) ||
        upload(h->g_switch, T.g_switch) || upload(h->g_atoms, T.g_atoms) || upload(h->g_nq, g_nq) ||
        upload(h->u_sw, T.u_sw) || upload(h->u_grp, T.u_grp) || upload(h->sp_off, T.sp_off) ||
        upload(h->sp_partner, T.sp_partner) || upload(h->gs_off, T.gs_off) || upload(h->gs_atoms, T.gs_atoms) || upload(h->nq_off, T.nq_off) || upload(h->nq_atoms, T.nq_atoms) ||
        upload(h->iqseq, T.iqseq0) || upload(h->is_q, T.is_q) || upload(h->excl, T.excl) ||
        upload(h->qbonded, T.qbonded) || upload(h->u_excl, T.u_excl) || upload(h->ljcode, T.ljcode) ||
        upload(h->sp_code, T.sp_code))
        return 1;
    {
        std::vector<QPar4> a(T.qp_tab.size()), b(T.qw_tab.size());
        for (size_t k = 0; k < a.size(); k++) a[k] = QPar4{T.qp_tab[k].A, T.qp_tab[k].B, T.qp_tab[k].el, T.qp_tab[k].score};
        for (size_t k = 0; k < b.size(); k++) b[k] = QPar4{T.qw_tab[k].A, T.qw_tab[k].B, T.qw_tab[k].el, T.qw_tab[k].score};
        if (upload(h->qp_tab, a) || upload(h->qw_tab, b)) return 1;
        std::vector<float4> af(a.size()), bf(b.size());
        for (size_t k = 0; k < a.size(); k++) af[k] = make_float4((float)a[k].A, (float)a[k].B, (float)a[k].el, (float)a[k].score);
        for (size_t k = 0; k < b.size(); k++) bf[k] = make_float4((float)b[k].A, (float)b[k].B, (float)b[k].el, (float)b[k].score);
        if (upload(h->qp_tabf, af) || upload(h->qw_tabf, bf)) return 1;
        std::vector<QStatic> qs;
        for (const auto &e : T.qq_list) qs.push_back(QStatic{e.i, e.j, e.state, e.soft, QPar4{e.p.A, e.p.B, e.p.el, e.p.score}});
        h->n_qq = (int)qs.size();
        for (const auto &e : T.qqp_list) qs.push_back(QStatic{e.i, e.j, e.state, 0, QPar4{e.p.A, e.p.B, e.p.el, e.p.score}});
        h->n_qstatic = (int)qs.size();
        if (upload(h->qstatic, qs)) return 1;
    }
    D.crg = h->crg.p; D.crgf = h->crgf.p; D.ctype = h->ctype.p; D.is_q = h->is_q.p; D.excl = h->excl.p; D.qbonded = h->qbonded.p;
    D.grp_of_atom = h->grp_of_atom.p; D.g_first = h->g_first.p; D.g_n = h->g_n.p; D.g_switch = h->g_switch.p;
    D.g_atoms = h->g_atoms.p; D.g_nq = h->g_nq.p; D.u_sw = h->u_sw.p; D.u_grp = h->u_grp.p; D.u_excl = h->u_excl.p;
    D.ljf = h->ljf.p; D.ljd = h->ljd.p; D.ljcode = h->ljcode.p; D.sp_off = h->sp_off.p; D.sp_partner = h->sp_partner.p; D.sp_code = h->sp_code.p;
    D.gs_off = h->gs_off.p; D.gs_atoms = h->gs_atoms.p; D.nq_off = h->nq_off.p; D.nq_atoms = h->nq_atoms.p; D.iqseq = h->iqseq.p; D.qp_tab = h->qp_tab.p; D.qw_tab = h->qw_tab.p; D.qp_tabf = h->qp_tabf.p; D.qw_tabf = h->qw_tabf.p;
    for (int a = 0; a < 3; a++) {
        D.wq[a] = s.nwat > 0 ? (float)T.w_crg[a] : 0.f;
        D.wqd[a] = s.nwat > 0 ? T.w_crg[a] : 0.0;
        D.wct[a] = s.nwat > 0 ? T.w_ctype[a] : 0;
        for (int b = 0; b < 3; b++) {
            const QPar p = s.nwat > 0 ? T.ww_par[a * 3 + b] : QPar{};
            D.wwA[a * 3 + b] = (float)p.A; D.wwB[a * 3 + b] = (float)p.B; D.wwQ[a * 3 + b] = (float)p.el; D.wwQd[a * 3 + b] = p.el; D.wwAd[a * 3 + b] = p.A; D.wwBd[a * 3 + b] = p.B;
        }
    }
    // ---- round-2 row kernels (qnb_rows.cuh): type-pair tables with the combination rule resolved, {12 A, 6 B} in FP32
    {
        const int nct = T.nct;
        if (nct > 255) return fail("qnb_init: %d atom types in use; the packed records hold the compact type in 8 bits", nct);
        std::vector<float2> ljp((size_t)2 * nct * nct), pw0((size_t)std::max(nct, 1));
        std::vector<float4> pw12((size_t)std::max(nct, 1));
        for (int a = 0; a < nct; a++)
            for (int b = 0; b < nct; b++)
                for (int plane = 0; plane < 2; plane++) {
                    const int code = plane ? 3 : T.ljcode[(size_t)a * nct + b];
                    double A, B;
                    T.combine(T.lj_a[a * 3 + code - 1], T.lj_b[a * 3 + code - 1], T.lj_a[b * 3 + code - 1], T.lj_b[b * 3 + code - 1], A, B);
                    ljp[((size_t)plane * nct + a) * nct + b] = make_float2((float)(12.0 * A), (float)(6.0 * B));
                }
        h->pw_hlj = false;
        RowPar &R = h->rowpar;
        EnergyPar &E = h->epar;
        if (s.nwat > 0) {
            double A[3], B[3];
            for (int a = 0; a < nct; a++) {
                for (int site = 0; site < 3; site++) {
                    const int code = T.ljcode[(size_t)a * nct + T.w_ctype[site]];
                    T.combine(T.lj_a[a * 3 + code - 1], T.lj_b[a * 3 + code - 1], T.lj_a[T.w_ctype[site] * 3 + code - 1],
                              T.lj_b[T.w_ctype[site] * 3 + code - 1], A[site], B[site]);
                }
                pw0[a] = make_float2((float)(12.0 * A[0]), (float)(6.0 * B[0]));
                pw12[a] = make_float4((float)(12.0 * A[1]), (float)(12.0 * A[2]), (float)(6.0 * B[1]), (float)(6.0 * B[2]));
                if (A[1] != 0.0 || A[2] != 0.0 || B[1] != 0.0 || B[2] != 0.0) h->pw_hlj = true;
            }
            // packed groups of k_water_rows: (a,b) = {(1,0),(2,0)} {(1,1),(2,2)} {(1,2),(2,1)} {(0,1),(0,2)}
            static const int ga[4][2] = {{1, 2}, {1, 2}, {1, 2}, {0, 0}}, gb[4][2] = {{0, 0}, {1, 2}, {2, 1}, {1, 2}};
            for (int g = 0; g < 4; g++) {
                const QPar &p0 = T.ww_par[ga[g][0] * 3 + gb[g][0]], &p1 = T.ww_par[ga[g][1] * 3 + gb[g][1]];
                R.wQ[g] = make_float2((float)p0.el, (float)p1.el);
                R.wA12[g] = make_float2((float)(12.0 * p0.A), (float)(12.0 * p1.A));
                R.wB6[g] = make_float2((float)(6.0 * p0.B), (float)(6.0 * p1.B));
            }
            R.q00 = (float)T.ww_par[0].el; R.A00 = (float)(12.0 * T.ww_par[0].A); R.B00 = (float)(6.0 * T.ww_par[0].B);
            R.wq0 = (float)T.w_crg[0]; R.wq12 = make_float2((float)T.w_crg[1], (float)T.w_crg[2]);
            for (int k = 0; k < 9; k++) { E.wwQ[k] = T.ww_par[k].el; E.wwA[k] = T.ww_par[k].A; E.wwB[k] = T.ww_par[k].B; }
            for (int k = 0; k < 3; k++) { E.wq[k] = T.w_crg[k]; E.wct[k] = T.w_ctype[k]; }
        }
        R.el14 = (float)s.el14_scale; R.nct = nct;
        E.el14 = s.el14_scale; E.nct = nct; E.geometric = D.geometric; E.any_atom = D.any_atom; E.spc = D.spc_water;
        if (upload(h->ljp, ljp) || upload(h->pw0, pw0) || upload(h->pw12, pw12)) return 1;
        // any-atom periodic lists carry the image of the pair's reference atoms, which need not be the minimum image
        // of the switching atoms the fixed-point rows work with: the round-1 general kernels keep that case
        h->legacy_rows = (D.any_atom && D.use_PBC) || getenv("QNB_LEGACY_ROWS");
    }
    const size_t n3 = 3 * (size_t)s.natom;
    h->nE = QNB_E_COUNT + QNB_EQ_STRIDE * s.nstates;
    h->nout = n3 + (size_t)kESlots * h->nE + kRstOut;   // [gradient | energy slots | restraint results]
    if (h->x.ensure(n3 + kMaxStates)) return 1;
    {
        // one allocation of whole 2 MiB pages for everything other ranks read or write (a cudaMalloc of this size is never
        // a sub-allocation of a driver slab, so the IPC handle of its base maps exactly these bytes in the peers)
        const size_t nlrf = (size_t)QNB_LRF_STRIDE * std::max(s.ncgp, 1);
        const size_t a16 = 16;   // doubles: keep every part 128-byte aligned
        h->arena_lrf_off = (h->nout + a16 - 1) / a16 * a16;
        h->arena_ctl_off = h->arena_lrf_off + (nlrf + a16 - 1) / a16 * a16;
        const size_t bytes = ((h->arena_ctl_off + 512) * sizeof(double) + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20);
        cudaError_t e = cudaMalloc(&h->arena, bytes);
        if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
        h->arena_doubles = bytes / sizeof(double);
        CU(cudaMemset(h->arena, 0, bytes));
        h->out.alias(h->arena, h->nout);
        h->lrf.alias(h->arena + h->arena_lrf_off, nlrf);
    }
    CU(cudaMallocHost(&h->hx, (n3 + kMaxStates) * sizeof(double)));
    h->hlam = h->hx + n3;
    for (int k = 0; k < kMaxStates; k++) h->hlam[k] = 0.0;
    h->lam_dev = h->x.p + n3;
    CU(cudaMallocHost(&h->hout, h->nout * sizeof(double)));
    {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, h->device));
        h->nsm = prop.multiProcessorCount;
    }
    {
        // The main stream carries the chain of short list-build kernels with its host round trips (and the step's pack
        // kernel): most urgent.  The LRF accumulation (1.5 ms on C5, thousands of blocks) runs next to that chain on a stream
        // of its own at the lowest priority, so that the chain's blocks are placed first whenever an SM slot frees up
        // (until r04 it ran on aux[4], a most-urgent stream, and the chain queued behind it).
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const bool flat = getenv("QNB_NO_PRIORITY") != nullptr;
        CU(cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, flat || getenv("QNB_ST_LOW") ? lo : hi));
        CU(cudaStreamCreateWithPriority(&h->lrf_st, cudaStreamNonBlocking, flat || !getenv("QNB_LRF_HIGH") ? lo : hi));
        CU(cudaEventCreateWithFlags(&h->ev_lrf, cudaEventDisableTiming));
    }
    for (int k = 0; k < kAux; k++) {
    {
        // the longest kernel of a step is placed first: stream k carries kernel k (kStreamOf); the solute rows (stream 1)
        // are the critical path of a solvated protein, the water rows fill the SMs behind them
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = most urgent
        static const int kRank[kAux] = {1, 0, 1, 0, 0, 1};   // water, solute, q_partner, q_atom, the tiny kernels (static lists, restraints), energies
        int pr = std::min(lo, hi + (getenv("QNB_NO_PRIORITY") ? 0 : kRank[k]));
        CU(cudaStreamCreateWithPriority(&h->aux[k], cudaStreamNonBlocking, pr));
    }
        CU(cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    for (int k = 0; k < 6; k++) CU(cudaEventCreate(&h->ev_bt[k]));
    if (const char *e = getenv("QNB_NO_HOST_REGISTER")) h->host_direct = !(e[0] == '1');
    if (const char *e = getenv("QNB_NO_GRAPH")) h->use_graph = !(e[0] == '1');
    if (const char *e = getenv("QNB_ONE_STREAM")) h->multi_stream = !(e[0] == '1');
    if (const char *e = getenv("QNB_WATER_BLOCKS")) h->water_blocks = std::max(0, atoi(e));
    if (const char *e = getenv("QNB_SOLUTE_BLOCKS")) h->solute_blocks = std::max(0, atoi(e));
    return 0;
}

// Sum of a part of the arena over the ranks, in place on every rank: the kernel over peer memory when the ranks have
// attached each other's arenas (graph-capturable: a plain kernel), else NCCL.
static bool p2p_ready(const qnb_handle *h) { return h->p2p.n > 1; }
static int allreduce_arena(qnb_handle *h, size_t off, size_t count, cudaStream_t st) {
    if (p2p_ready(h)) {
        // blocks for this rank's slice (1/n of the elements), two 16-byte elements per thread at most
        const size_t slice2 = ((count + 1) / 2 + h->p2p.n - 1) / h->p2p.n;
        const int blocks = std::max(1, std::min(h->nsm, (int)((slice2 + 511) / 512)));
        LAUNCH_ON(h, st, k_p2p_allreduce, blocks, 256, 0, h->p2p, off, count,
                  reinterpret_cast<unsigned *>(h->arena + h->arena_ctl_off));
        return 0;
    }
    if (!h->comm) return 0;
    int rc = g_nccl.AllReduce(h->arena + off, h->arena + off, count, kNcclDouble, kNcclSum, h->comm, st);
    if (rc) return fail("ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return 0;
}

// ------------------------------------------------------------------ list build
// A list build (or a box change) alters launch arguments baked into the captured step.  The executable graphs are
// kept and refreshed with cudaGraphExecUpdate from a new capture (tens of microseconds) instead of being
// re-instantiated (hundreds).
static void drop_graphs(qnb_handle *h, bool destroy = false) {
    h->build_serial++;
    for (int c = 0; c < 2; c++)
        for (int f = 0; f < 16; f++) {
            h->graph_dirty[c][f] = true;
            if (destroy && h->graph[c][f]) { cudaGraphExecDestroy(h->graph[c][f]); h->graph[c][f] = nullptr; }
        }
}

static void make_grid(qnb_handle *h, const double *hx) {
    const qnb_system &s = h->T.s;
    Grid &G = h->grid;
    // any-atom mode: the deciding atoms may sit up to rmax from either switch atom
    double rmax2 = 0.0;
    if (h->D.any_atom) {
        double r2m = 0.0;
        for (int g = 0; g < s.ncgp_solute; g++) {
            const int sw = h->T.g_switch[g];
            for (int k = 0; k < h->T.g_n[g]; k++) {
                const int a = h->T.g_atoms[h->T.g_first[g] + k];
                double d2 = 0;
                for (int d = 0; d < 3; d++) d2 += (hx[3 * a + d] - hx[3 * sw + d]) * (hx[3 * a + d] - hx[3 * sw + d]);
                r2m = std::max(r2m, d2);
            }
        }
        rmax2 = 2.0 * std::sqrt(r2m) + 1e-6;
    }
    h->cut.rmax2 = rmax2;
    const double rcmax = (std::sqrt(std::max({h->cut.rc2[0], h->cut.rc2[1], h->cut.rc2[2], 1.0})) + rmax2) * 1.0001;
    // cells of HALF the largest cut-off, searched with reach 2: the scanned volume is (5/2 Rc)^3 instead of (3 Rc)^3
    // for the lists and hugs the LRF shell twice as tightly (env QNB_CELL_DIV overrides the divisor)
    int cell_div = 2;
    if (const char *e = getenv("QNB_CELL_DIV")) cell_div = std::max(1, atoi(e));
    const double edge = rcmax / cell_div;
    const int nmax = 48;
    if (s.use_PBC) {
        G.periodic = 1;
        for (int d = 0; d < 3; d++) {
            int n = (int)std::floor(h->box[d] / edge);
            n = std::max(1, std::min(n, nmax));
            G.n[d] = n;
            G.org[d] = 0;
            G.inv_box[d] = h->inv_box[d];
            G.inv_cell[d] = n / h->box[d];
        }
    } else {
        G.periodic = 0;
        // bounding box of the switch atoms at the first build, padded; later drift is absorbed by clamping
        if (!h->have_grid) {
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int u = 0; u < h->T.nunit; u++)
                for (int d = 0; d < 3; d++) {
                    double v = hx[3 * h->T.u_sw[u] + d];
                    lo[d] = std::min(lo[d], v); hi[d] = std::max(hi[d], v);
                }
            for (int d = 0; d < 3; d++) { G.org[d] = lo[d] - 2.0; G.inv_box[d] = hi[d] + 2.0 - G.org[d]; /* extent kept here */ }
        }
        for (int d = 0; d < 3; d++) {
            const double ext = G.inv_box[d];
            double cell = std::max(edge, ext / nmax);
            G.n[d] = std::max(1, (int)std::floor(ext / cell) + 1);
            G.inv_cell[d] = 1.0 / cell;
        }
    }
    G.ncell = G.n[0] * G.n[1] * G.n[2];
    h->have_grid = true;
    // LRF reach in cells
    const bool all = h->cut.lrf_all[0] || h->cut.lrf_all[1] || h->cut.lrf_all[2] ||
                     (h->D.any_atom && s.use_PBC && h->cut.rclrf2 == -1.0);
    const double rl = std::sqrt(std::max(h->cut.rclrf2, 0.0)) + rmax2;
    int r[3];
    for (int d = 0; d < 3; d++) {
        double m = std::ceil(rl * G.inv_cell[d] * 1.0001);   // |cell index difference| <= ceil(R/edge), also after clamping
        r[d] = (all || m > G.n[d]) ? G.n[d] : (int)m;
    }
    h->lrf_reach = make_int3(r[0], r[1], r[2]);
    for (int d = 0; d < 3; d++) {
        double m = std::ceil(rcmax * G.inv_cell[d]);
        r[d] = std::max(1, (int)m);
    }
    h->list_reach = make_int3(r[0], r[1], r[2]);
}

static int run_exclusive_scan(qnb_handle *h, const int *in, int *out, int n) {
    LAUNCH(h, k_exclusive_scan, 1, 1024, 0, in, out, n);
    return 0;
}

// nbqplist / nbqwlist: flags -> scan -> compaction, the counts copied back asynchronously (the caller synchronises)
static int launch_q_lists(qnb_handle *h) {
    const Dev &D = h->D;
    if (D.nqat <= 0) return 0;
    const double rex2 = h->T.s.rexcl_o * h->T.s.rexcl_o;
    const bool skip_qp = h->qp_done && (D.use_PBC ? (h->cut.Rq < 0.0) : (h->cut.rcq2 > rex2));
    const bool skip_qw = h->qw_done && (D.use_PBC ? (h->cut.Rq < 0.0) : (h->cut.rcq2 > rex2));
    if (!skip_qp && D.ncgp_solute > 0) {
        LAUNCH(h, k_qp_flags, cdiv(D.ncgp_solute, 128), 128, 0, D, h->cut, h->x.p, h->flag.p, h->qp_shift_atom.p);
        LAUNCH(h, k_exclusive_scan, 1, 1024, 0, h->flag.p, h->pos.p, D.nat_solute);
        LAUNCH(h, k_compact, cdiv(D.nat_solute, 256), 256, 0, D.nat_solute, h->flag.p, h->pos.p, h->qp_list.p);
        CU(cudaMemcpyAsync(&h->nqp, h->pos.p + D.nat_solute, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        h->qp_done = true;
    }
    if (!skip_qw && D.nwat > 0) {
        LAUNCH(h, k_qw_flags, cdiv(D.nwat, 128), 128, 0, D, h->cut, h->x.p, h->flag2.p);
        LAUNCH(h, k_exclusive_scan, 1, 1024, 0, h->flag2.p, h->pos2.p, D.nwat);
        LAUNCH(h, k_compact, cdiv(D.nwat, 256), 256, 0, D.nwat, h->flag2.p, h->pos2.p, h->qw_list.p);
        CU(cudaMemcpyAsync(&h->nqw, h->pos2.p + D.nwat, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        h->qw_done = true;
    }
    return 0;
}

static int build_device(qnb_handle *h, const double *hx_for_grid) {
    const Dev &D = h->D;
    drop_graphs(h);   // row pointers, counts and list sizes are baked into the captured launches
    const int nu = D.nunit;
    if (nu >= (1 << 26)) return fail("qnb_build_lists: %d charge groups + waters, the LRF queue entries hold 26 bits of item index", nu);
    if (!h->restoring) {   // what qnb_save_lists would have to remember
        h->hx_build.assign(hx_for_grid, hx_for_grid + 3 * (size_t)D.natom);
        h->cut_build = h->cut;
    }
    make_grid(h, hx_for_grid);
    const Grid G = h->grid;
    refresh_frame(h);
    if (h->upos.ensure(3 * (size_t)std::max(nu, 1)) || h->cell_of.ensure(std::max(nu, 1)) ||
        h->cell_count.ensure(G.ncell + 1) || h->cell_start.ensure(G.ncell + 2) || h->cell_items.ensure(std::max(nu, 1)) ||
        h->counts.ensure(3 * (size_t)std::max(nu, 1)) || h->row_tot.ensure(std::max(nu, 1) + 1) ||
        h->row_off.ensure(std::max(nu, 1) + 2) ||
        h->flag.ensure(std::max({D.natom, D.nwat, 1}) + 1) || h->pos.ensure(std::max({D.natom, D.nwat, 1}) + 2) ||
        h->flag2.ensure(std::max(D.nwat, 1) + 1) || h->pos2.ensure(std::max(D.nwat, 1) + 2) ||
        h->qp_list.ensure(std::max(D.nat_solute, 1)) || h->qw_list.ensure(std::max(D.nwat, 1)) ||
        h->qp_shift_atom.ensure(std::max(D.ncgp_solute, 1)) || h->item_pos.ensure(std::max(nu, 1)) || h->item_posf.ensure(std::max(nu, 1)) || h->item_scr.ensure(std::max(nu, 1)) ||
        h->src.ensure(std::max(D.natom, 1)) || h->srcf.ensure(std::max(D.natom, 1)) || h->cell_unsorted.ensure(std::max(nu, 1)) ||
        h->lrf_planes.ensure((size_t)cdiv(std::max(nu, 1), 32) * kLrfPlaneBytes) ||
        h->item_nq.ensure(std::max(nu, 1) + 1) || h->src_off.ensure(std::max(nu, 1) + 2) ||
        h->pk_atom.ensure(D.natom + 4) || h->pk_ct.ensure(D.natom + 4) || h->pk_q.ensure(D.natom + 4) ||
        h->pk_qd.ensure(D.natom + 4) || h->px.ensure(D.natom + 4) || h->py.ensure(D.natom + 4) ||   // +4: the force kernels always fetch three sites
        h->pz.ensure(D.natom + 4) || h->upk.ensure(std::max(nu, 1)) || h->pk_sw.ensure(D.natom + 4) ||
        h->rec_i.ensure(D.natom + 4) || h->rec_f.ensure(D.natom + 4) || h->wT.ensure(D.natom + 4) || h->wown.ensure(3 * (size_t)std::max(D.nwat, 1)) || h->wd.ensure(2 * (size_t)(D.natom + 4)) ||
        h->e_cnt.ensure(std::max(nu, 1) + D.ncgp_solute + 4) || h->e_off.ensure(std::max(nu, 1) + D.ncgp_solute + 8))
        return 1;
    // LRF: cgp_centers + lrf_update over every pair that falls in the LRF branch.  It needs the cell tables and the
    // packed atoms but not the rows, and it is the longest kernel of a build: it runs on a side stream next to the
    // row/chunk kernels (which are short, latency-bound and interrupted by the host round trip for the sizes).
    bool lrf_forked = false;
    auto launch_lrf = [&]() {
        if (!(D.use_LRF && D.ncgp > 0) || h->restoring) return;
        cudaStream_t ls = h->multi_stream ? h->lrf_st : h->st;
        if (h->multi_stream) {
            cudaEventRecord(h->ev_fork, h->st);
            cudaStreamWaitEvent(ls, h->ev_fork, 0);
        }
        if (h->time_build) cudaEventRecord(h->ev_bt[0], ls);
        LAUNCH_ON(h, ls, k_cgp_centers, cdiv(D.ncgp, 128), 128, 0, D, h->x.p, h->lrf.p);
        if (nu > 0) {
            LAUNCH_ON(h, ls, k_pack_sources, cdiv(D.natom, 256), 256, 0, h->src_off.p + nu, h->pk_atom.p, h->x.p, h->crg.p, h->src.p, h->srcf.p);
            // compaction pays when the LRF shell holds a small part of the scanned cells (boxes, short RcLRF)
            const double rl = std::sqrt(std::max(h->cut.rclrf2, 0.0));
            const bool all = h->cut.lrf_all[0] || h->cut.lrf_all[1] || h->cut.lrf_all[2];
            double ext = 0;
            for (int d = 0; d < 3; d++) ext = std::max(ext, G.n[d] / G.inv_cell[d]);
            const bool compact = !all && rl < 0.9 * ext;
            // periodic image by cell row when the reach m leaves room in every dimension (|delta| < box/2 for every scanned cell needs 2(m+1) <= n)
            bool rowshift = G.periodic && !all;
            const int rr[3] = {h->lrf_reach.x, h->lrf_reach.y, h->lrf_reach.z};
            // One cell more than the scan itself needs (2(m+1) <= n): the kernel takes the pair's shift from the image of
            // the scanned row, which equals nint((x_switch - cgp_cent)/L) as long as the target's group centre lies within
            // the spare cell(s) of its switch atom - two cell edges with this margin.
            for (int d = 0; d < 3; d++) rowshift = rowshift && 2 * (rr[d] + 2) <= G.n[d];
            const bool general = D.any_atom || D.sharded;
            // sphere, switching-atom lists, unsharded, every cell within the LRF reach: all-pairs tiling
            int gmax = 0;
            for (int gn : h->T.g_n) gmax = std::max(gmax, gn);
            bool allpairs = !G.periodic && !general && gmax <= kLrfTileAtoms && !getenv("QNB_LRF_GENERAL");
            for (int d = 0; d < 3; d++) allpairs = allpairs && rr[d] >= G.n[d];
            if (allpairs && !h->lrf_mom.ensure((size_t)kLrfRaw * nu)) {
                cudaMemsetAsync(h->lrf_mom.p, 0, sizeof(double) * kLrfRaw * (size_t)nu, ls);
                const int tb = cdiv(nu, 128);
                // one resident wave of equal slices (r02o: 896 blocks on 740 slots = 1.2 waves, the second a fifth full)
                static const int ap_minb = [] { const char *e = getenv("QNB_LRFAP_MINB"); return e && atoi(e) == 5 ? 5 : 6; }();
                static int occ_ap = 0;
                if (!occ_ap) {
                    if (ap_minb == 5) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ap, k_lrf_allpairs<5>, 128, 0);
                    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ap, k_lrf_allpairs<6>, 128, 0);
                    occ_ap = std::max(occ_ap, 1);
                }
                // QNB_LRFAP_BLOCKS=n (< occupancy): leave SM room for the row scan and chunk kernels that run next to it (unused
                // dynamic shared memory caps the resident blocks)
                static const int ap_blocks = [] { const char *e = getenv("QNB_LRFAP_BLOCKS"); return e ? std::max(1, atoi(e)) : 0; }();
                const int nres = ap_blocks > 0 ? std::min(ap_blocks, occ_ap) : occ_ap;
                const size_t pad = nres < occ_ap ? (size_t)(228 * 1024 / (nres + 1) + 1024 > 17 * 1024 ? 228 * 1024 / (nres + 1) + 1024 - 17 * 1024 : 0) : 0;
                const int slices = std::max(1, std::min(cdiv(nu, kLrfTileItems), (nres * h->nsm) / tb));
                if (ap_minb == 5)
                    LAUNCH_ON(h, ls, k_lrf_allpairs<5>, dim3(tb, slices), 128, std::min<size_t>(pad, 30 * 1024), D, h->cut, h->x.p, h->upos.p, h->item_pos.p, h->item_posf.p,
                              h->src_off.p, h->src.p, h->lrf.p, h->lrf_mom.p);
                else
                    LAUNCH_ON(h, ls, k_lrf_allpairs<6>, dim3(tb, slices), 128, 0, D, h->cut, h->x.p, h->upos.p, h->item_pos.p, h->item_posf.p,
                              h->src_off.p, h->src.p, h->lrf.p, h->lrf_mom.p);
                LAUNCH_ON(h, ls, k_lrf_expand, cdiv(nu * 40, 256), 256, 0, D, h->lrf_mom.p, h->lrf.p);
            } else {
                LAUNCH_ON(h, ls, k_pack_lrf_planes, cdiv(nu, 128), 128, 0, D, nu, h->item_pos.p, h->src_off.p, h->src.p, h->lrf_planes.p);
#define LRFCASE(CP, RS, GN) LAUNCH_ON(h, ls, (k_lrf_accumulate<CP, RS, GN>), cdiv(nu, kRowWarps), 32 * kRowWarps, 0, D, h->cut, G, h->lrf_reach, h->x.p, h->upos.p, \
                                      h->cell_of.p, h->cell_start.p, h->cell_items.p, h->item_pos.p, h->item_posf.p, h->src_off.p, h->src.p, h->srcf.p, h->lrf_planes.p, h->lrf.p)
            if (rowshift) { if (general) LRFCASE(true, true, true); else LRFCASE(true, true, false); }
            else if (compact) { if (general) LRFCASE(true, false, true); else LRFCASE(true, false, false); }
            else { if (general) LRFCASE(false, false, true); else LRFCASE(false, false, false); }
#undef LRFCASE
            }
        }
        if (h->time_build) cudaEventRecord(h->ev_bt[1], ls);
        if (h->multi_stream) { cudaEventRecord(h->ev_lrf, ls); lrf_forked = true; }
    };
    const bool md_lists = true;
    if (nu > 0 && md_lists) {
        CU(cudaMemsetAsync(h->cell_count.p, 0, sizeof(int) * (G.ncell + 1), h->st));
        LAUNCH(h, k_bin_units, cdiv(nu, 256), 256, 0, D, G, h->x.p, h->upos.p, h->cell_of.p, h->cell_count.p);
        run_exclusive_scan(h, h->cell_count.p, h->cell_start.p, G.ncell);
        CU(cudaMemsetAsync(h->cell_count.p, 0, sizeof(int) * (G.ncell + 1), h->st));
        LAUNCH(h, k_cell_fill, cdiv(nu, 256), 256, 0, nu, h->cell_of.p, h->cell_start.p, h->cell_count.p, h->cell_unsorted.p);
        LAUNCH(h, k_cell_sort, G.ncell, 64, 0, G.ncell, h->cell_start.p, h->cell_unsorted.p, h->cell_items.p);
        LAUNCH(h, k_pack_items, cdiv(nu, 256), 256, 0, D, h->upos.p, h->cell_items.p, h->item_pos.p, h->item_posf.p, h->item_scr.p, h->item_nq.p);
        run_exclusive_scan(h, h->item_nq.p, h->src_off.p, nu);
        LAUNCH(h, k_pack_atoms, cdiv(nu, 128), 128, 0, D, h->cell_items.p, h->src_off.p, h->pk_atom.p, h->pk_q.p, h->pk_qd.p, h->pk_ct.p, h->pk_sw.p, h->upk.p);
        launch_lrf();
        // QNB_OLD_ROWS=1: the round-1 builder (one cell-row segment per step), kept for comparison
        static const bool old_rows = getenv("QNB_OLD_ROWS") != nullptr;
        if (h->time_build) cudaEventRecord(h->ev_bt[2], h->st);
        if (old_rows)
            LAUNCH(h, k_build_rows<false>, cdiv(nu * 32, 256), 256, 0, D, h->cut, G, h->list_reach, h->x.p, h->upos.p, h->cell_of.p, h->cell_start.p,
                   h->item_pos.p, h->src_off.p, h->counts.p, (const int *)nullptr, (uint32_t *)nullptr);
        else
            LAUNCH(h, k_rows_scan<false>, cdiv(nu, kRowScanWarps), 32 * kRowScanWarps, 0, D, h->cut, G, h->list_reach, h->x.p, h->upos.p, h->cell_of.p,
                   h->cell_start.p, h->item_pos.p, h->item_scr.p, h->src_off.p, h->counts.p, (const int *)nullptr, (uint32_t *)nullptr);
        if (h->time_build) cudaEventRecord(h->ev_bt[3], h->st);
        LAUNCH(h, k_row_totals, cdiv(nu, 256), 256, 0, nu, h->counts.p, h->row_tot.p);
        // chunk tables of the streaming force kernels and flat pair lists of the energy kernel: counts now, contents after
        // the rows are filled; all eight prefix sums in one launch
        const int nsol = D.ncgp_solute, nwat = D.nwat;
        if (h->nch.ensure(nu + 2) || h->choff.ensure(nu + 4) || h->ucost.ensure(nu + 2) || h->cost_off.ensure(nu + 4)) return 1;
        int *nch_w = h->nch.p, *nch_s = h->nch.p + nwat + 1, *off_w = h->choff.p, *off_s = h->choff.p + nwat + 2;
        int *uc_w = h->ucost.p, *uc_s = h->ucost.p + nwat + 1, *co_w = h->cost_off.p, *co_s = h->cost_off.p + nwat + 2;
        int *en_ww = h->e_cnt.p, *en_pp = h->e_cnt.p + nwat, *en_pw = h->e_cnt.p + nwat + nsol;
        int *eo_ww = h->e_off.p, *eo_pp = h->e_off.p + nwat + 1, *eo_pw = h->e_off.p + nwat + nsol + 2;
        const bool new_rows = !h->legacy_rows;
        if (nwat > 0) LAUNCH(h, k_chunk_count, cdiv(nwat, 256), 256, 0, D, nsol, nwat, 0, new_rows, h->counts.p, nch_w, uc_w);
        if (nsol > 0) LAUNCH(h, k_chunk_count, cdiv(nsol, 256), 256, 0, D, 0, nsol, kITile, new_rows, h->counts.p, nch_s, uc_s);
        LAUNCH(h, k_energy_counts, cdiv(nu, 256), 256, 0, D, h->counts.p, en_ww, en_pp, en_pw);
        {
            ScanJobs J{};
            int nj = 0;
            auto job = [&](const int *in, int *out, int n) { J.in[nj] = in; J.out[nj] = out; J.n[nj] = n; nj++; };
            job(h->row_tot.p, h->row_off.p, nu);
            job(nch_w, off_w, nwat); job(uc_w, co_w, nwat); job(nch_s, off_s, nsol); job(uc_s, co_s, nsol);
            job(en_ww, eo_ww, nwat); job(en_pp, eo_pp, nsol); job(en_pw, eo_pw, nsol);
            LAUNCH(h, k_multi_scan, nj, 1024, 0, J);
        }
        int total = 0, npk = 0;
        h->nwchunk = h->nschunk = 0;
        CU(cudaMemcpyAsync(&total, h->row_off.p + nu, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(&npk, h->src_off.p + nu, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(&h->nwchunk, off_w + nwat, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(&h->nschunk, off_s + nsol, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(&h->n_ww_e, eo_ww + nwat, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(&h->n_pp_e, eo_pp + nsol, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(&h->n_pw_e, eo_pw + nsol, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        // Q-atom partner lists: built once when the cut-off covers everything (nbqplist L3678, nbqwlist L3889,
        // nbqplist_box L3780, nbqwlist_box L3972); their counts ride on the same host round trip as the row sizes
        if (launch_q_lists(h)) return 1;
        CU(cudaStreamSynchronize(h->st));   // the only host round trip of the build: sizes for the allocations
        h->total_rows = total;
        h->npk = npk;
        if (h->rows.ensure((size_t)std::max(total, 1)) || h->wdesc.ensure(std::max(h->nwchunk, 1)) ||
            h->wrow.ensure((size_t)std::max(h->nwchunk, 1) * 32) || h->sdesc.ensure(std::max(h->nschunk, 1)) ||
            h->srow.ensure((size_t)std::max(h->nschunk, 1) * 32) || h->sspec.ensure((size_t)std::max(h->nschunk, 1) * 32) ||
            h->ww_pairs.ensure(std::max(h->n_ww_e, 1)) || h->pp_pairs.ensure(std::max(h->n_pp_e, 1)) || h->pw_pairs.ensure(std::max(h->n_pw_e, 1)))
            return 1;
        if (h->time_build) cudaEventRecord(h->ev_bt[4], h->st);
        if (old_rows)
            LAUNCH(h, k_build_rows<true>, cdiv(nu * 32, 256), 256, 0, D, h->cut, G, h->list_reach, h->x.p, h->upos.p, h->cell_of.p, h->cell_start.p,
                   h->item_pos.p, h->src_off.p, h->counts.p, h->row_off.p, h->rows.p);
        else
            LAUNCH(h, k_rows_scan<true>, cdiv(nu, kRowScanWarps), 32 * kRowScanWarps, 0, D, h->cut, G, h->list_reach, h->x.p, h->upos.p, h->cell_of.p,
                   h->cell_start.p, h->item_pos.p, h->item_scr.p, h->src_off.p, h->counts.p, h->row_off.p, h->rows.p);
        if (h->time_build) cudaEventRecord(h->ev_bt[5], h->st);
        if (h->n_ww_e + h->n_pp_e + h->n_pw_e > 0)
            LAUNCH(h, k_energy_fill, cdiv(nu * 32, 256), 256, 0, D, h->counts.p, h->row_off.p, h->rows.p, h->upk.p, h->pk_atom.p, eo_ww, eo_pp,
                   eo_pw, h->ww_pairs.p, h->pp_pairs.p, h->pw_pairs.p);
        // the all-zero records behind the last packed atom: what the padding lanes of the water-row chunks load
        CU(cudaMemsetAsync(h->wT.p + npk, 0, 3 * sizeof(float4), h->st));
        CU(cudaMemsetAsync(h->rec_i.p + npk, 0, sizeof(int4), h->st));
        CU(cudaMemsetAsync(h->rec_f.p + npk, 0, sizeof(float4), h->st));
        if (h->nwchunk > 0)
            LAUNCH(h, k_chunk_fill, cdiv(nwat * 32, 256), 256, 0, D, nsol, nwat, 0, h->counts.p, h->row_off.p, h->rows.p, off_w,
                   h->wdesc.p, h->wrow.p, h->pk_atom.p, (uint16_t *)nullptr, new_rows, (uint32_t)npk);
        if (h->nschunk > 0)
            LAUNCH(h, k_chunk_fill, cdiv(nsol * 32, 256), 256, 0, D, 0, nsol, kITile, h->counts.p, h->row_off.p, h->rows.p, off_s,
                   h->sdesc.p, h->srow.p, h->pk_atom.p, h->sspec.p, false, 0u);
        // one resident wave per kernel, every warp an equal share of the estimated work
        const int occw = new_rows ? h->occ_wr : h->occ_w, occs = new_rows ? h->occ_sr : h->occ_s;
        const int min_chunks = new_rows ? 2 : 4;   // chunks per warp below which more blocks only add launch overhead
        const int sh = std::max(1, h->share);
        h->wgrid = std::max(1, std::min(cdiv((h->water_blocks ? h->water_blocks : std::max(occw, 1)) * h->nsm, sh), cdiv(h->nwchunk, 4 * min_chunks)));
        h->sgrid = std::max(1, std::min(cdiv((h->solute_blocks ? h->solute_blocks : std::max(occs, 1)) * h->nsm, sh), cdiv(h->nschunk, 4 * min_chunks)));
        if (h->wstart_w.ensure(4 * h->wgrid + 2) || h->wstart_s.ensure(4 * h->sgrid + 2)) return 1;
        if (h->nwchunk > 0)
            LAUNCH(h, k_warp_starts, cdiv(4 * h->wgrid + 1, 128), 128, 0, D, nsol, nwat, 0, new_rows, h->counts.p, off_w, co_w, 4 * h->wgrid, h->wstart_w.p);
        if (h->nschunk > 0)
            LAUNCH(h, k_warp_starts, cdiv(4 * h->sgrid + 1, 128), 128, 0, D, 0, nsol, kITile, new_rows, h->counts.p, off_s, co_s, 4 * h->sgrid, h->wstart_s.p);
    }
    if (!(nu > 0 && md_lists)) { if (launch_q_lists(h)) return 1; }
    // LRF: the moments were started on the side stream after the cell tables (see above); join, then the exchange
    if (D.use_LRF && D.ncgp > 0 && !h->restoring) {
        if (lrf_forked) CU(cudaStreamWaitEvent(h->st, h->ev_lrf, 0));
        if (h->comm || p2p_ready(h)) {
            // lrf_gather (nonbondene.f90:616-623): sum the moments over the ranks; the centres (identical on every rank)
            // are summed along and recomputed afterwards: bit-identical to the single-rank value
            if (allreduce_arena(h, h->arena_lrf_off, (size_t)QNB_LRF_STRIDE * D.ncgp, h->st)) return 1;
            LAUNCH(h, k_cgp_centers_only, cdiv(D.ncgp, 128), 128, 0, D, h->x.p, h->lrf.p);
        }
    }
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    h->lists_built = true;
    return 0;
}

// ------------------------------------------------------------------ one nonbonded evaluation on device
enum StepKernel { K_WATER = 0, K_SOLUTE, K_QPARTNER, K_QATOM, K_QSTATIC, K_LRF, K_RST, K_ENERGY, K_COUNT };
static const char *kStepKernelNames[K_COUNT] = {"k_water_rows", "k_solute_rows", "k_q_partner", "k_q_atom",
                                                "k_qq_static", "k_lrf_taylor", "k_solvent_restraints", "k_pair_energy"};

constexpr int kFlagEnergiesOnly = 64;   // internal (QCP beads): no kernel whose only product is a gradient

static bool step_kernel_active(const qnb_handle *h, int k, int flags) {
    const Dev &D = h->D;
    const bool md = flags & QNB_FLAG_MD;
    switch (k) {
    case K_WATER: return md && D.nwat > 0 && h->nwchunk > 0;
    case K_SOLUTE: return md && D.ncgp_solute > 0 && h->nschunk > 0;
    case K_QPARTNER: return !(flags & kFlagEnergiesOnly) && D.nqat > 0 && (h->nqp + h->nqw) > 0;
    case K_QATOM: return D.nqat > 0 && (h->nqp + h->nqw) > 0;
    case K_QSTATIC: return (flags & QNB_FLAG_QQ) && h->T.s.is_master && h->n_qstatic > 0;
    case K_LRF: return md && D.use_LRF;
    // restrain_solvent / watpol: only in_md and only for the sphere (potene.f90:161-167)
    // (one rank only when the step is sharded: every rank would add the same term to the summed gradient)
    case K_RST: return md && (flags & QNB_FLAG_SOLVENT_RESTRAINTS) && h->rst_set && !D.use_PBC && D.nwat > 0 && h->T.s.is_master;
    // E%pp, E%pw, E%ww: every listed pair once, FP64 (the row kernels produce gradients only)
    case K_ENERGY: return md && !(flags & QNB_FLAG_NO_ENERGY) && (h->n_ww_e + h->n_pp_e + h->n_pw_e) > 0;
    }
    return false;
}

static size_t solute_smem(const Dev &D) { return (size_t)D.nct * 6 * (sizeof(double) + sizeof(float)) + (size_t)D.nct * D.nct; }

// resident blocks per SM of the two persistent kernels in the variant this system uses (register-limited)
static void query_occupancy(qnb_handle *h) {
    const Dev &D = h->D;
    const bool pbc = D.use_PBC, spc = D.spc_water, geom = D.geometric;
    const size_t sm = solute_smem(D);
#define WOCC(P, S, G) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_w, k_water_force<P, S, G>, 128, 0)
    if (pbc) { if (spc) WOCC(true, true, true); else if (geom) WOCC(true, false, true); else WOCC(true, false, false); }
    else { if (spc) WOCC(false, true, true); else if (geom) WOCC(false, false, true); else WOCC(false, false, false); }
#undef WOCC
#define SOCC(P, G) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_s, k_solute_force<P, G>, 128, sm)
    if (pbc) { if (geom) SOCC(true, true); else SOCC(true, false); }
    else { if (geom) SOCC(false, true); else SOCC(false, false); }
#undef SOCC
    if (const char *e = getenv("QNB_WROWS_MINB")) h->wrows_minb = std::min(6, std::max(4, atoi(e)));
    if (h->wrows_minb == 4) {
        if (spc) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_wr, k_water_rows<true, 4>, 128, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_wr, k_water_rows<false, 4>, 128, 0);
    } else if (h->wrows_minb == 5) {
        if (spc) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_wr, k_water_rows<true, 5>, 128, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_wr, k_water_rows<false, 5>, 128, 0);
    } else {
        if (spc) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_wr, k_water_rows<true, 6>, 128, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_wr, k_water_rows<false, 6>, 128, 0);
    }
    if (h->pw_hlj) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_sr, k_solute_rows<true>, 128, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_sr, k_solute_rows<false>, 128, 0);
    cudaGetLastError();
}

// grids of the Q and energy kernels (shared by the single launches and the fused step kernels)
static dim3 q_partner_grid(const qnb_handle *h) {
    const int n = h->nqp + 3 * h->nqw;
    return dim3(cdiv(n, 128), std::max(1, std::min(h->D.nqat, cdiv(cdiv(6 * 148, std::max(1, h->share)), cdiv(n, 128)))));
}
static dim3 q_atom_grid(const qnb_handle *h) {
    static const int qa_mult = [] { const char *e = getenv("QNB_QATOM_BLOCKS_PER_SM"); return e ? std::max(1, atoi(e)) : 4; }();
    const int nsite = h->nqp + 3 * h->nqw;
    return dim3(h->D.nqat, std::max(1, std::min(cdiv(nsite, 128), cdiv(cdiv(qa_mult * 148, std::max(1, h->share)), std::max(h->D.nqat, 1)))));
}
static int energy_grid(const qnb_handle *h) {
    const int n = h->n_ww_e + h->n_pp_e + h->n_pw_e;
    return std::max(1, std::min(cdiv(n, 128), cdiv(8 * h->nsm, std::max(1, h->share))));
}

static void launch_step_kernel(qnb_handle *h, int k, cudaStream_t cs, int flags) {
    const Dev &D = h->D;
    double *grad = h->out.p, *E = h->out.p + 3 * (size_t)D.natom;
    const int nE = h->nE;
    const bool pbc = D.use_PBC, spc = D.spc_water, geom = D.geometric;
    // the round-1 general kernels serve any-atom periodic lists (QNB_LEGACY_ROWS=1 forces them), gradient only:
    // the energies always come from k_pair_energy
    const int want_e = 0;
    switch (k) {
    case K_WATER: {
        if (!h->legacy_rows) {
#define WROWS(S, M)                                                                                                            \
    do {                                                                                                                       \
        using B_ = WaterRowsBody<S, M>;                                                                                        \
        STEP_LAUNCH(h, cs, B_, 128, M, (k_water_rows<S, M>), h->wgrid, 0, h->rowpar, h->rec_i.p, h->rec_f.p, h->wT.p, h->wown.p, \
                    h->pw0.p, h->pw12.p, h->wstart_w.p, h->wdesc.p, h->wrow.p, D.nat_solute, grad);                           \
    } while (0)
            if (h->wrows_minb == 4) { if (spc) WROWS(true, 4); else WROWS(false, 4); }
            else if (h->wrows_minb == 5) { if (spc) WROWS(true, 5); else WROWS(false, 5); }
            else { if (spc) WROWS(true, 6); else WROWS(false, 6); }
#undef WROWS
            break;
        }
#define WCASE(P, S, G)                                                                                                             \
    LAUNCH_ON(h, cs, (k_water_force<P, S, G>), h->wgrid, 128, 0, D, h->x.p, h->px.p, h->py.p, h->pz.p, h->pk_q.p, h->pk_ct.p,     \
              h->pk_atom.p, h->wstart_w.p, h->wdesc.p, h->wrow.p, grad, E, nE, want_e)
        if (pbc) { if (spc) WCASE(true, true, true); else if (geom) WCASE(true, false, true); else WCASE(true, false, false); }
        else { if (spc) WCASE(false, true, true); else if (geom) WCASE(false, false, true); else WCASE(false, false, false); }
#undef WCASE
        break;
    }
    case K_SOLUTE: {
        if (!h->legacy_rows) {
            if (h->pw_hlj) STEP_LAUNCH(h, cs, SoluteRowsBody<true>, 128, QNB_SROWS_MINB, k_solute_rows<true>, h->sgrid, 0, h->rowpar, h->upk.p, h->nq_off.p, h->rec_i.p, h->rec_f.p, h->wT.p,
                                     h->ljp.p, h->pw0.p, h->pw12.p, h->wstart_s.p, h->sdesc.p, h->srow.p, h->sspec.p, h->pk_atom.p, grad);
            else STEP_LAUNCH(h, cs, SoluteRowsBody<false>, 128, QNB_SROWS_MINB, k_solute_rows<false>, h->sgrid, 0, h->rowpar, h->upk.p, h->nq_off.p, h->rec_i.p, h->rec_f.p, h->wT.p,
                           h->ljp.p, h->pw0.p, h->pw12.p, h->wstart_s.p, h->sdesc.p, h->srow.p, h->sspec.p, h->pk_atom.p, grad);
            break;
        }
        const size_t sm = solute_smem(D);
#define SCASE(P, G)                                                                                                              \
    LAUNCH_ON(h, cs, (k_solute_force<P, G>), h->sgrid, 128, sm, D, h->x.p, h->px.p, h->py.p, h->pz.p, h->pk_q.p, h->pk_qd.p,    \
              h->pk_ct.p, h->pk_atom.p, h->wstart_s.p, h->sdesc.p, h->srow.p, h->sspec.p, grad, E, nE, want_e)
        if (pbc) { if (geom) SCASE(true, true); else SCASE(true, false); }
        else { if (geom) SCASE(false, true); else SCASE(false, false); }
#undef SCASE
        break;
    }
    case K_QPARTNER: {
        const size_t sm = sizeof(float) * (3 * (size_t)D.nqat + D.nstates);
        const dim3 pgrid = q_partner_grid(h);
        if (pbc) STEP_LAUNCH(h, cs, QPartnerBody<true>, 128, 6, k_q_partner<true>, pgrid, sm, D, h->x.p, h->lam_dev, h->nqp, h->qp_list.p, h->qp_shift_atom.p, h->nqw, h->qw_list.p, grad);
        else STEP_LAUNCH(h, cs, QPartnerBody<false>, 128, 6, k_q_partner<false>, pgrid, sm, D, h->x.p, h->lam_dev, h->nqp, h->qp_list.p, h->qp_shift_atom.p, h->nqw, h->qw_list.p, grad);
        break;
    }
    case K_QATOM: {
        const dim3 qgrid = q_atom_grid(h);
#define QCASE(P, N)                                                                                                           \
    do {                                                                                                                      \
        using B_ = QAtomBody<P, N>;                                                                                           \
        STEP_LAUNCH(h, cs, B_, 128, 5, (k_q_atom<P, N>), qgrid, 0, D, h->x.p, h->lam_dev, h->nqp, h->qp_list.p, h->qp_shift_atom.p, \
                    h->nqw, h->qw_list.p, grad, E, nE);                                                                       \
    } while (0)
        const int ns = D.nstates;
        if (pbc) { if (ns <= 1) QCASE(true, 1); else if (ns <= 2) QCASE(true, 2); else if (ns <= 4) QCASE(true, 4); else QCASE(true, 8); }
        else { if (ns <= 1) QCASE(false, 1); else if (ns <= 2) QCASE(false, 2); else if (ns <= 4) QCASE(false, 4); else QCASE(false, 8); }
#undef QCASE
        break;
    }
    case K_QSTATIC:
        STEP_LAUNCH(h, cs, QqStaticBody, 128, 6, k_qq_static, cdiv(h->n_qstatic, 128), 0, h->n_qstatic, h->n_qq, h->qstatic.p, h->x.p, h->lam_dev, grad, E, nE);
        break;
    case K_LRF:
        STEP_LAUNCH(h, cs, LrfTaylorBody, 128, 6, k_lrf_taylor, cdiv(D.natom, 128), 0, D, h->x.p, h->lrf.p, grad, E, nE);
        break;
    case K_ENERGY: {
        const int grid = energy_grid(h);
        if (pbc) STEP_LAUNCH(h, cs, PairEnergyBody<true>, 128, 7, k_pair_energy<true>, grid, 0, h->epar, h->n_ww_e, h->ww_pairs.p, h->n_pp_e, h->pp_pairs.p, h->n_pw_e,
                           h->pw_pairs.p, h->px.p, h->py.p, h->pz.p, h->pk_qd.p, h->pk_ct.p, h->pk_sw.p, h->x.p, h->wd.p, h->ljd.p, h->ljcode.p, E, nE);
        else STEP_LAUNCH(h, cs, PairEnergyBody<false>, 128, 7, k_pair_energy<false>, grid, 0, h->epar, h->n_ww_e, h->ww_pairs.p, h->n_pp_e, h->pp_pairs.p, h->n_pw_e,
                       h->pw_pairs.p, h->px.p, h->py.p, h->pz.p, h->pk_qd.p, h->pk_ct.p, h->pk_sw.p, h->x.p, h->wd.p, h->ljd.p, h->ljcode.p, E, nE);
        break;
    }
    case K_RST: {
        const qnb_solvent_restraints &r = h->rst;
        RstPar P{};
        for (int c = 0; c < 3; c++) P.xw[c] = r.xwcent[c];
        P.rsurf = r.rwat - r.shift; P.fk = r.fk_wsphere; P.Dwmz = r.Dwmz; P.awmz = r.awmz; P.fkwpol = r.fkwpol;
        P.nsh = r.wpol_restr ? r.nwpolr_shell : 0;
        P.wpol = P.nsh > 0;
        for (int k = 0; k < P.nsh; k++) { P.rout[k] = r.rout[k]; P.cstb[k] = r.cstb[k]; P.tcorr[k] = h->theta_corr[k]; }
        if (P.nsh > 0) P.rin_last = r.rout[P.nsh - 1] - r.dr[P.nsh - 1];
        double *rst = h->out.p + 3 * (size_t)D.natom + (size_t)kESlots * nE;
        int *shell_cnt = reinterpret_cast<int *>(rst + 2 + 2 * QNB_MAX_SHELLS);   // zeroed with the output buffer
        LAUNCH_ON(h, cs, k_rst_theta, cdiv(D.nwat, 128), 128, 0, D, P, h->x.p, grad, rst, h->wp_theta.p, shell_cnt, h->wp_shell_list.p, h->wp_shell_theta.p);
        if (P.wpol)
            LAUNCH_ON(h, cs, k_rst_watpol, dim3(std::min(cdiv(D.nwat, 4), 192), P.nsh), 128, 0, D, P, h->x.p, grad, rst, h->wp_theta.p,
                      shell_cnt, h->wp_shell_list.p, h->wp_shell_theta.p);
        break;
    }
    }
}

// ---- the fused step: the row kernels in one launch, everything else in a second one (k_uber, qnb_kernels.cuh)
using URowsSolute = SoluteRowsBody<false>;
using URowsWater = WaterRowsBody<true, 5>;
static bool uber_ok(const qnb_handle *h, int flags) {
    // Opt-in (QNB_UBER=1).  Measured r03g: the step takes the same time with three launches as with nine (C2 32.8 vs 32.4 us,
    // C3 22.5, C5 87.1) -- the graph's node count is not what bounds it; the two row kernels each need the whole machine for
    // their single wave (solute 10.3 + water 11.6 us) and follow the pack kernel (4.5 us).
    static const bool on = getenv("QNB_UBER") != nullptr;
    const Dev &D = h->D;
    return on && !h->collect && h->multi_stream && !h->legacy_rows && (D.spc_water || D.nwat == 0) && !h->pw_hlj && D.nstates <= 2 &&
           !(flags & QNB_FLAG_SOLVENT_RESTRAINTS);
}
template <bool PBC, int NS>
static void launch_uber_rest(qnb_handle *h, cudaStream_t cs, int flags) {
    const Dev &D = h->D;
    double *grad = h->out.p, *E = h->out.p + 3 * (size_t)D.natom;
    USlot<QAtomBody<PBC, NS>> qa{};
    USlot<QPartnerBody<PBC>> qp{};
    USlot<PairEnergyBody<PBC>> en{};
    USlot<LrfTaylorBody> lt{};
    USlot<QqStaticBody> qs{};
    size_t sm = 0;
    if (step_kernel_active(h, K_QATOM, flags)) {
        const dim3 g = q_atom_grid(h);
        qa.a = body_args_t<QAtomBody<PBC, NS>>(D, h->x.p, h->lam_dev, h->nqp, h->qp_list.p, h->qp_shift_atom.p, h->nqw, h->qw_list.p, grad, E, h->nE);
        qa.gx = (int)g.x; qa.gy = (int)g.y;
    }
    if (step_kernel_active(h, K_QPARTNER, flags)) {
        const dim3 g = q_partner_grid(h);
        qp.a = body_args_t<QPartnerBody<PBC>>(D, h->x.p, h->lam_dev, h->nqp, h->qp_list.p, h->qp_shift_atom.p, h->nqw, h->qw_list.p, grad);
        qp.gx = (int)g.x; qp.gy = (int)g.y;
        sm = sizeof(float) * (3 * (size_t)D.nqat + D.nstates);
    }
    if (step_kernel_active(h, K_ENERGY, flags)) {
        en.a = body_args_t<PairEnergyBody<PBC>>(h->epar, h->n_ww_e, h->ww_pairs.p, h->n_pp_e, h->pp_pairs.p, h->n_pw_e, h->pw_pairs.p, h->px.p, h->py.p,
                                                h->pz.p, h->pk_qd.p, h->pk_ct.p, h->pk_sw.p, h->x.p, h->wd.p, h->ljd.p, h->ljcode.p, E, h->nE);
        en.gx = energy_grid(h); en.gy = 1;
    }
    if (step_kernel_active(h, K_LRF, flags)) {
        lt.a = body_args_t<LrfTaylorBody>(D, h->x.p, h->lrf.p, grad, E, h->nE);
        lt.gx = cdiv(D.natom, 128); lt.gy = 1;
    }
    if (step_kernel_active(h, K_QSTATIC, flags)) {
        qs.a = body_args_t<QqStaticBody>(h->n_qstatic, h->n_qq, h->qstatic.p, h->x.p, h->lam_dev, grad, E, h->nE);
        qs.gx = cdiv(h->n_qstatic, 128); qs.gy = 1;
    }
    const int nb = qa.gx * qa.gy + qp.gx * qp.gy + en.gx * en.gy + lt.gx * lt.gy + qs.gx * qs.gy;
    if (nb <= 0) return;
    k_uber<6, QAtomBody<PBC, NS>, QPartnerBody<PBC>, PairEnergyBody<PBC>, LrfTaylorBody, QqStaticBody><<<nb, 128, sm, cs>>>(qa, qp, en, lt, qs);
    h->launches++;
}
static void launch_uber_rows(qnb_handle *h, cudaStream_t cs, int flags) {
    const Dev &D = h->D;
    double *grad = h->out.p;
    USlot<URowsSolute> so{};
    USlot<URowsWater> wa{};
    if (step_kernel_active(h, K_SOLUTE, flags)) {
        so.a = body_args_t<URowsSolute>(h->rowpar, h->upk.p, h->nq_off.p, h->rec_i.p, h->rec_f.p, h->wT.p, h->ljp.p, h->pw0.p, h->pw12.p, h->wstart_s.p,
                                        h->sdesc.p, h->srow.p, h->sspec.p, h->pk_atom.p, grad);
        so.gx = h->sgrid; so.gy = 1;
    }
    if (step_kernel_active(h, K_WATER, flags)) {
        wa.a = body_args_t<URowsWater>(h->rowpar, h->rec_i.p, h->rec_f.p, h->wT.p, h->wown.p, h->pw0.p, h->pw12.p, h->wstart_w.p, h->wdesc.p, h->wrow.p,
                                       D.nat_solute, grad);
        wa.gx = h->wgrid; wa.gy = 1;
    }
    const int nb = so.gx + wa.gx;
    if (nb <= 0) return;
    k_uber<4, URowsSolute, URowsWater><<<nb, 128, 0, cs>>>(so, wa);
    h->launches++;
}

// The kernels of one evaluation are independent (they only meet in atomicAdd on grad/E), so they are issued on
// four streams between a fork and a join event: at 12k atoms no single kernel fills 148 SMs.
static const int kStreamOf[K_COUNT] = {0, 1, 2, 3, 4, -1, 4, 5};   // aux stream index, -1 = main stream

static int issue_step(qnb_handle *h, int flags, bool out_cleared = false) {
    // the output buffer is cleared by the packing kernel when there is one (one graph node less on the critical path)
    const bool pack = (flags & QNB_FLAG_MD) && h->npk > 0;
    if (!out_cleared && !pack) CU(cudaMemsetAsync(h->out.p, 0, h->nout * sizeof(double), h->st));
    if (pack)
        LAUNCH(h, k_pack_step, cdiv(h->npk, 256), 256, 0, h->npk, h->fix, h->D.nat_solute, h->pk_atom.p, h->pk_sw.p, h->pk_q.p, h->pk_ct.p,
               h->x.p, h->px.p, h->py.p, h->pz.p, h->rec_i.p, h->rec_f.p, h->wT.p, h->wown.p, h->wd.p, out_cleared ? (double *)nullptr : h->out.p,
               (int)h->nout, (const double *)nullptr, (double *)nullptr, 0);
    CU(cudaEventRecord(h->ev_fork, h->st));
    bool used[kAux] = {};
    if (uber_ok(h, flags)) {
        // two launches instead of seven: [solute rows | water rows] and [Q atoms | Q partners | energies | LRF Taylor | static Q lists]
        const bool rows = step_kernel_active(h, K_SOLUTE, flags) || step_kernel_active(h, K_WATER, flags);
        if (rows) {
            CU(cudaStreamWaitEvent(h->aux[1], h->ev_fork, 0)); used[1] = true;
            launch_uber_rows(h, h->aux[1], flags);
        }
        CU(cudaStreamWaitEvent(h->aux[3], h->ev_fork, 0)); used[3] = true;
        const bool pbc = h->D.use_PBC;
        if (h->D.nstates <= 1) { if (pbc) launch_uber_rest<true, 1>(h, h->aux[3], flags); else launch_uber_rest<false, 1>(h, h->aux[3], flags); }
        else { if (pbc) launch_uber_rest<true, 2>(h, h->aux[3], flags); else launch_uber_rest<false, 2>(h, h->aux[3], flags); }
        for (int k = 0; k < kAux; k++)
            if (used[k]) {
                CU(cudaEventRecord(h->ev_join[k], h->aux[k]));
                CU(cudaStreamWaitEvent(h->st, h->ev_join[k], 0));
            }
        return 0;
    }
    static const int kOrder[K_COUNT] = {K_RST, K_QSTATIC, K_SOLUTE, K_QATOM, K_QPARTNER, K_WATER, K_ENERGY, K_LRF};
    for (int o = 0; o < K_COUNT; o++) {
        const int k = kOrder[o];
        if (!step_kernel_active(h, k, flags)) continue;
        const int si = h->multi_stream ? kStreamOf[k] : -1;
        cudaStream_t cs = si < 0 ? h->st : h->aux[si];
        if (si >= 0 && !used[si]) { CU(cudaStreamWaitEvent(cs, h->ev_fork, 0)); used[si] = true; }
        launch_step_kernel(h, k, cs, flags);
    }
    for (int k = 0; k < kAux; k++)
        if (used[k]) {
            CU(cudaEventRecord(h->ev_join[k], h->aux[k]));
            CU(cudaStreamWaitEvent(h->st, h->ev_join[k], 0));
        }
    return 0;
}

// d(host, mapped) += gradient: the last node of an end-to-end step when the caller's d is page-locked
__global__ void __launch_bounds__(256) k_add_out(size_t n, const double *__restrict__ grad, double *__restrict__ d_host) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(d_host) & 15u) == 0) {
        const size_t n2 = n / 2;
        if (i < n2) {
            double2 v = reinterpret_cast<double2 *>(d_host)[i];
            const double2 g = reinterpret_cast<const double2 *>(grad)[i];
            v.x += g.x; v.y += g.y;
            reinterpret_cast<double2 *>(d_host)[i] = v;
        }
        if (i == 0 && (n & 1)) d_host[n - 1] += grad[n - 1];
    } else {
        for (size_t k = 2 * i; k < 2 * i + 2 && k < n; k++) d_host[k] += grad[k];
    }
}

// Device alias of a caller buffer that qnb_register_host_buffers page-locked (kept until qnb_finalize /
// qnb_release_host_buffers); nullptr for any other buffer (the call then stages through the handle's own pinned memory).
static void *host_alias(qnb_handle *h, const void *p, size_t bytes, bool may_register = false) {
    if (!h->host_direct || !p) return nullptr;
    for (const auto &r : h->hostregs)
        if (r.p == p && (r.bytes >= bytes || r.bytes == 0)) return r.dev;
    if (!may_register || h->hostregs.size() >= 32) return nullptr;
    qnb_handle::HostReg r{p, bytes, nullptr};
    cudaError_t e = cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e == cudaSuccess || e == cudaErrorHostMemoryAlreadyRegistered) {
        if (e != cudaSuccess) cudaGetLastError();
        void *dev = nullptr;
        if (cudaHostGetDevicePointer(&dev, const_cast<void *>(p), 0) == cudaSuccess) r.dev = dev;
        else { cudaGetLastError(); if (e == cudaSuccess) cudaHostUnregister(const_cast<void *>(p)); }
        if (e != cudaSuccess && r.dev) r.bytes = 0;   // somebody else's registration: never unregistered here
    } else cudaGetLastError();
    h->hostregs.push_back(r);
    return r.dev;
}
static void release_host_buffers(qnb_handle *h) {
    for (const auto &r : h->hostregs)
        if (r.dev && r.bytes) cudaHostUnregister(const_cast<void *>(r.p));
    cudaGetLastError();
    h->hostregs.clear();
    h->x_direct = nullptr; h->d_direct = nullptr;
}

// One evaluation on the device.  with_copies: pinned x -> device before, [grad|E|EQ] -> pinned after.
// Captured once per list build into a CUDA graph (one launch per MD step instead of ~10 API calls).
static int step_device(qnb_handle *h, int flags, bool with_copies = false, cudaGraph_t *graph_out = nullptr, int *nodes_out = nullptr) {
    const size_t n3 = 3 * (size_t)h->T.s.natom;
    // sharded step: the all-reduce is issued eagerly after the kernels unless QNB_SHARD_GRAPH=1 asks for it to be captured
    // with them (NCCL >= 2.9 supports capture once the communicator has been used; not yet run on hardware)
    static const bool shard_graph = getenv("QNB_SHARD_GRAPH") != nullptr;
    const bool graphable = h->use_graph && (!h->comm || shard_graph || p2p_ready(h));
    flags &= (QNB_FLAG_MD | QNB_FLAG_QQ | QNB_FLAG_NO_ENERGY | QNB_FLAG_SOLVENT_RESTRAINTS);
    const int gi = (flags & 7) | ((flags & QNB_FLAG_SOLVENT_RESTRAINTS) ? 8 : 0);
    // the copies of an end-to-end step: x from the caller's page-locked array (else from the staging copy), lambda from
    // the staging tail; afterwards either [grad|E|EQ] -> pinned (added on the host) or d += grad by the device and only
    // the energy tail -> pinned
    auto copy_in = [&]() -> int {
        int rc = 0;
        if (h->x_direct) {
            rc |= cudaMemcpyAsync(h->x.p, h->x_direct, n3 * sizeof(double), cudaMemcpyHostToDevice, h->st) != cudaSuccess;
            rc |= cudaMemcpyAsync(h->lam_dev, h->hlam, kMaxStates * sizeof(double), cudaMemcpyHostToDevice, h->st) != cudaSuccess;
        } else
            rc |= cudaMemcpyAsync(h->x.p, h->hx, (n3 + kMaxStates) * sizeof(double), cudaMemcpyHostToDevice, h->st) != cudaSuccess;
        return rc;
    };
    auto copy_out = [&]() -> int {
        int rc = 0;
        if (h->d_direct) {
            if (h->d_overwrite) rc |= cudaMemcpyAsync(h->d_host, h->out.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st) != cudaSuccess;
            else LAUNCH(h, k_add_out, cdiv((int)((n3 + 1) / 2), 256), 256, 0, n3, h->out.p, h->d_direct);
            rc |= cudaMemcpyAsync(h->hout + n3, h->out.p + n3, (h->nout - n3) * sizeof(double), cudaMemcpyDeviceToHost, h->st) != cudaSuccess;
        } else
            rc |= cudaMemcpyAsync(h->hout, h->out.p, h->nout * sizeof(double), cudaMemcpyDeviceToHost, h->st) != cudaSuccess;
        return rc;
    };
    if (graphable) {
        cudaGraphExec_t &ge = h->graph[with_copies ? 1 : 0][gi];
        bool &dirty = h->graph_dirty[with_copies ? 1 : 0][gi];
        if (with_copies && (h->graph_x[gi] != h->x_direct || h->graph_d[gi] != h->d_direct || h->graph_dow[gi] != h->d_overwrite)) dirty = true;   // baked pointers
        if (!ge || dirty || graph_out) {
            cudaGraph_t g = nullptr;
            const int64_t l0 = h->launches;
            CU(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
            int rc = 0;
            bool cleared = false;
            if (with_copies) {
                if (h->multi_stream) {
                    // clearing the output does not depend on the upload: a parallel branch of the graph
                    rc |= cudaEventRecord(h->ev_pack, h->st) != cudaSuccess;
                    rc |= cudaStreamWaitEvent(h->aux[4], h->ev_pack, 0) != cudaSuccess;
                    rc |= cudaMemsetAsync(h->out.p, 0, h->nout * sizeof(double), h->aux[4]) != cudaSuccess;
                    rc |= cudaEventRecord(h->ev_join[4], h->aux[4]) != cudaSuccess;
                    cleared = true;
                }
                rc |= copy_in();
                if (cleared) rc |= cudaStreamWaitEvent(h->st, h->ev_join[4], 0) != cudaSuccess;
            }
            rc |= issue_step(h, flags, cleared);
            rc |= allreduce_arena(h, 0, h->nout, h->st);
            if (with_copies) { rc |= copy_out(); h->graph_x[gi] = h->x_direct; h->graph_d[gi] = h->d_direct; h->graph_dow[gi] = h->d_overwrite; }
            cudaError_t ce = cudaStreamEndCapture(h->st, &g);
            h->graph_launches[with_copies ? 1 : 0][gi] = (int)(h->launches - l0);
            h->launches = l0;
            if (rc || ce != cudaSuccess || !g) return fail("CUDA graph capture of the step failed: %s", cudaGetErrorString(ce));
            if (graph_out) {   // the caller (batched step) makes this graph a child of its own; this handle's own graph stays as it was
                *graph_out = g;
                if (nodes_out) *nodes_out = h->graph_launches[with_copies ? 1 : 0][gi];
                dirty = true;   // the pointers recorded above belong to the caller's graph now
                return 0;
            }
            bool updated = false;
            if (ge) {
                cudaGraphExecUpdateResultInfo info;
                updated = cudaGraphExecUpdate(ge, g, &info) == cudaSuccess;
                if (!updated) { cudaGetLastError(); cudaGraphExecDestroy(ge); ge = nullptr; }
            }
            if (!updated) ce = cudaGraphInstantiate(&ge, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) return fail("cudaGraphInstantiate: %s", cudaGetErrorString(ce));
            dirty = false;
        }
        CU(cudaGraphLaunch(ge, h->st));
        h->launches += h->graph_launches[with_copies ? 1 : 0][gi];
        return 0;
    }
    if (graph_out) return fail("internal: this handle's step cannot be captured into a graph");
    if (with_copies && copy_in()) return fail("step: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (issue_step(h, flags)) return 1;
    // gather_nonbond + serial sum on the master (potene.f90:195-222) as one all-reduce over [d | E | EQ]
    if (allreduce_arena(h, 0, h->nout, h->st)) return 1;
    if (with_copies && copy_out()) return fail("step: download failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

}  // namespace qnb

namespace qnb {
template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T *out, int iters, T a, T b) {
    T v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (T)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = v[k] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += v[k];
    if (s == (T)123456789) out[0] = s;   // never true: keeps the chain alive
}

}  // namespace qnb

// =============================================================================== C ABI
extern "C" {

const char *qnb_last_error(void) { return g_err.c_str(); }

#ifdef QNB_TRACE
int qnb_trace_read(unsigned long long *out) {   // [2][8192][10], experiment builds only
    CU(cudaMemcpyFromSymbol(out, qnb::g_trace, sizeof(unsigned long long) * 2 * 8192 * 10));
    return 0;
}
int qnb_trace_clear(void) {
    static std::vector<unsigned long long> z(2 * 8192 * 10, 0ull);
    CU(cudaMemcpyToSymbol(qnb::g_trace, z.data(), sizeof(unsigned long long) * z.size()));
    return 0;
}
#endif

int qnb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { fail("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return 0; }
    return n;
}

int qnb_init(const qnb_system *sys, int device, qnb_handle **out) {
    if (!sys || !out) return fail("qnb_init: null argument");
    if (sys->abi_version != QNB_ABI_VERSION) return fail("qnb_init: abi_version %d, library is %d", sys->abi_version, QNB_ABI_VERSION);
    int n = qnb_device_count();
    if (n <= 0) return fail("qnb_init: no CUDA device (%s); this library has no CPU path", g_err.c_str());
    if (device < 0 || device >= n) return fail("qnb_init: device %d out of range (0..%d)", device, n - 1);
    CU(cudaSetDevice(device));
    qnb_handle *h = new qnb_handle();
    h->device = device;
    if (!h->T.build(sys)) { fail("qnb_init: %s", h->T.error.c_str()); delete h; return 1; }
    if (init_device(h)) { delete h; return 1; }
    qnb::query_occupancy(h);
    *out = h;
    return 0;
}

int qnb_update_box(qnb_handle *h, const double boxlength[3], const double inv_boxl[3]) {
    if (!h) return fail("null handle");
    for (int d = 0; d < 3; d++) { h->box[d] = boxlength[d]; h->inv_box[d] = inv_boxl[d]; }
    refresh_dev(h);
    drop_graphs(h);
    return 0;
}

int qnb_set_solvent_restraints(qnb_handle *h, const qnb_solvent_restraints *p) {
    if (!h || !p) return fail("null argument");
    if (p->nwpolr_shell < 0 || p->nwpolr_shell > QNB_MAX_SHELLS) return fail("qnb_set_solvent_restraints: %d shells (max %d)", (int)p->nwpolr_shell, QNB_MAX_SHELLS);
    if (h->T.s.use_PBC) return fail("qnb_set_solvent_restraints: restrain_solvent/watpol belong to the spherical boundary (potene.f90:161)");
    CU(cudaSetDevice(h->device));
    const int nw = std::max(h->T.s.nwat, 1);
    if (h->wp_theta.ensure(nw) || h->wp_shell_n.ensure(QNB_MAX_SHELLS) || h->wp_shell_list.ensure((size_t)nw * QNB_MAX_SHELLS) ||
        h->wp_shell_theta.ensure((size_t)nw * QNB_MAX_SHELLS)) return 1;
    h->rst = *p;
    h->rst_set = true;
    drop_graphs(h);   // the parameters are baked into the captured launches
    return 0;
}

int qnb_set_theta_corr(qnb_handle *h, const double *theta_corr) {
    if (!h || !theta_corr) return fail("null argument");
    if (!h->rst_set) return fail("qnb_set_theta_corr: call qnb_set_solvent_restraints first");
    for (int k = 0; k < h->rst.nwpolr_shell; k++) h->theta_corr[k] = theta_corr[k];
    drop_graphs(h);
    return 0;
}

int qnb_last_restraints(qnb_handle *h, double E[2], double *shell_theta_sum, int32_t *shell_n) {
    if (!h || !E || !shell_theta_sum || !shell_n) return fail("null argument");
    if (!h->rst_set) return fail("qnb_last_restraints: no solvent restraints set");
    E[0] = h->last_rst[0]; E[1] = h->last_rst[1];
    for (int k = 0; k < h->rst.nwpolr_shell; k++) {
        shell_theta_sum[k] = h->last_rst[2 + k];
        shell_n[k] = (int32_t)h->last_rst[2 + QNB_MAX_SHELLS + k];
    }
    return 0;
}

int qnb_set_constraints(qnb_handle *h, int nmol, const int32_t *mol_first, const int32_t *ij, const double *dist2,
                        const double *winv) {
    if (!h || !mol_first || !ij || !dist2 || !winv) return fail("qnb_set_constraints: null argument");
    if (nmol < 0) return fail("qnb_set_constraints: nmol < 0");
    CU(cudaSetDevice(h->device));
    const int natom = h->T.s.natom;
    if (mol_first[0] != 0) return fail("qnb_set_constraints: mol_first[0] must be 0");
    const int nc = nmol > 0 ? mol_first[nmol] : 0;
    std::vector<int> first(mol_first, mol_first + nmol + 1);
    std::vector<ShakePair> cij((size_t)std::max(nc, 0));
    std::vector<int> owner((size_t)natom, -1);   // molecules must not share atoms: one thread relaxes one molecule
    for (int m = 0; m < nmol; m++) {
        const int n = first[m + 1] - first[m];
        if (n < 0) return fail("qnb_set_constraints: mol_first must not decrease");
        if (n > kMaxMolConstraints)
            return fail("qnb_set_constraints: molecule %d has %d constraints; the device SHAKE handles solvent-sized molecules "
                        "(<= %d constraints), shake_solute stays on the host", m + 1, n, kMaxMolConstraints);
        for (int c = first[m]; c < first[m + 1]; c++) {
            const int i = ij[2 * c], j = ij[2 * c + 1];
            if (i < 1 || i > natom || j < 1 || j > natom || i == j) return fail("qnb_set_constraints: constraint %d: atoms %d %d out of range", c + 1, i, j);
            if (!(dist2[c] > 0.0)) return fail("qnb_set_constraints: constraint %d: dist2 must be positive", c + 1);
            for (int a : {i - 1, j - 1}) {
                if (owner[a] >= 0 && owner[a] != m) return fail("qnb_set_constraints: atom %d is constrained in two molecules", a + 1);
                owner[a] = m;
            }
            cij[c] = ShakePair{i - 1, j - 1};
        }
    }
    std::vector<double> d2(dist2, dist2 + std::max(nc, 0)), w(winv, winv + natom);
    if (upload(h->shk_first, first) || upload(h->shk_ij, cij) || upload(h->shk_d2, d2) || upload(h->shk_winv, w)) return 1;
    const size_t n3 = 3 * (size_t)natom;
    if (h->shk_x.ensure(n3) || h->shk_xx.ensure(n3) || h->shk_iter.ensure(2)) return 1;
    h->shk_nmol = nmol;
    h->shk_set = true;
    return 0;
}

int qnb_shake(qnb_handle *h, const double *xx, double *x, int64_t *iterations) {
    if (!h || !x) return fail("qnb_shake: null argument");
    if (!h->shk_set) return fail("qnb_shake: no constraints (call qnb_set_constraints)");
    if (!xx && !h->x_from_nonbond)
        return fail("qnb_shake: xx == NULL needs the coordinates of this step's qnb_nonbond resident on the device "
                    "(qnb_build_lists, qnb_restore_lists and qnb_qcp_beads replace them)");
    CU(cudaSetDevice(h->device));
    const size_t n3 = 3 * (size_t)h->T.s.natom;
    // xx == NULL: the reference coordinates are the ones already resident from this step's qnb_nonbond / qnb_build_lists
    const double *dxx = h->x.p;
    if (xx) {
        memcpy(h->hout, xx, n3 * sizeof(double));   // pinned staging (free between two nonbonded calls)
        CU(cudaMemcpyAsync(h->shk_xx.p, h->hout, n3 * sizeof(double), cudaMemcpyHostToDevice, h->st));
        dxx = h->shk_xx.p;
    }
    memcpy(h->hx, x, n3 * sizeof(double));
    CU(cudaMemcpyAsync(h->shk_x.p, h->hx, n3 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CU(cudaMemsetAsync(h->shk_iter.p, 0, 2 * sizeof(unsigned long long), h->st));
    if (h->shk_nmol > 0)
        LAUNCH(h, k_shake, cdiv(h->shk_nmol, 128), 128, 0, h->shk_nmol, h->shk_first.p, h->shk_ij.p, h->shk_d2.p, h->shk_winv.p,
               dxx, h->shk_x.p, h->shk_iter.p, (int *)(h->shk_iter.p + 1));
    // hout is reused for the way back: the copies are ordered on h->st
    unsigned long long res[2] = {0, 0};
    CU(cudaMemcpyAsync(h->hout, h->shk_x.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(res, h->shk_iter.p, sizeof res, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    if ((int)(res[1] & 0xffffffffu)) return fail("shake failure");   // bondene.f90:1143
    memcpy(x, h->hout, n3 * sizeof(double));
    if (iterations) *iterations = (int64_t)res[0];
    return 0;
}

int qnb_save_lists(qnb_handle *h) {
    if (!h) return fail("null handle");
    if (!h->lists_built) return fail("qnb_save_lists: pair lists have not been built");
    CU(cudaSetDevice(h->device));
    const size_t nl = (size_t)QNB_LRF_STRIDE * std::max(h->T.s.ncgp, 1);
    if (h->lrf_saved.ensure(nl)) return 1;
    CU(cudaMemcpyAsync(h->lrf_saved.p, h->lrf.p, nl * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->hx_saved = h->hx_build;
    h->cut_saved = h->cut_build;
    for (int d = 0; d < 3; d++) { h->box_saved[d] = h->box[d]; h->inv_saved[d] = h->inv_box[d]; }
    h->have_saved = true;
    return 0;
}

int qnb_restore_lists(qnb_handle *h) {
    if (!h) return fail("null handle");
    if (!h->have_saved) return fail("qnb_restore_lists: nothing saved (call qnb_save_lists)");
    CU(cudaSetDevice(h->device));
    for (int d = 0; d < 3; d++) { h->box[d] = h->box_saved[d]; h->inv_box[d] = h->inv_saved[d]; }
    refresh_dev(h);
    h->cut = h->cut_saved;
    const size_t n3 = 3 * (size_t)h->T.s.natom;
    memcpy(h->hx, h->hx_saved.data(), n3 * sizeof(double));
    CU(cudaMemcpyAsync(h->x.p, h->hx, n3 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    h->restoring = true;
    h->x_from_nonbond = false;
    const int rc = build_device(h, h->hx_saved.data());
    h->restoring = false;
    if (rc) return 1;
    h->hx_build = h->hx_saved;
    h->cut_build = h->cut_saved;
    const size_t nl = (size_t)QNB_LRF_STRIDE * std::max(h->T.s.ncgp, 1);
    CU(cudaMemcpyAsync(h->lrf.p, h->lrf_saved.p, nl * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int qnb_qcp_beads(qnb_handle *h, const double *x_save, int natq, const int32_t *atoms, int nbeads, const double *coord,
                  const double *lambda, double *EQ_out) {
    if (!h || !x_save || !atoms || !coord || !lambda || !EQ_out) return fail("qnb_qcp_beads: null argument");
    if (!h->lists_built) return fail("qnb_qcp_beads: pair lists have not been built (call qnb_build_lists)");
    if (natq <= 0 || nbeads <= 0) return fail("qnb_qcp_beads: natq and nbeads must be positive");
    CU(cudaSetDevice(h->device));
    const qnb_system &s = h->T.s;
    const size_t n3 = 3 * (size_t)s.natom;
    const int nEQ = QNB_EQ_STRIDE * s.nstates;
    std::vector<int> a0(natq);
    std::vector<double> base(3 * (size_t)natq);
    for (int j = 0; j < natq; j++) {
        if (atoms[j] < 1 || atoms[j] > s.natom) return fail("qnb_qcp_beads: atom %d out of range", (int)atoms[j]);
        a0[j] = atoms[j] - 1;
        for (int c = 0; c < 3; c++) base[3 * j + c] = x_save[3 * (size_t)a0[j] + c];
    }
    h->x_from_nonbond = false;   // the last bead's displaced coordinates stay on the device
    DBuf<int> &d_atoms = h->bead_atoms;
    DBuf<double> &d_base = h->bead_base, &d_disp = h->bead_disp, &d_eq = h->bead_eq;
    if (upload(d_atoms, a0) || upload(d_base, base) || d_disp.ensure(3 * (size_t)natq * nbeads) || d_eq.ensure((size_t)h->nE * nbeads)) return 1;
    memcpy(h->hx, x_save, n3 * sizeof(double));
    for (int k = 0; k < s.nstates; k++) h->hlam[k] = lambda[k];
    CU(cudaMemcpyAsync(h->x.p, h->hx, n3 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CU(cudaMemcpyAsync(h->lam_dev, h->hlam, kMaxStates * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CU(cudaMemcpyAsync(d_disp.p, coord, sizeof(double) * 3 * (size_t)natq * nbeads, cudaMemcpyHostToDevice, h->st));
    for (int b = 0; b < nbeads; b++) {
        LAUNCH(h, k_set_bead, cdiv(3 * natq, 128), 128, 0, natq, d_atoms.p, d_base.p, d_disp.p + 3 * (size_t)natq * b, h->x.p);
        if (issue_step(h, QNB_FLAG_QQ | kFlagEnergiesOnly)) return 1;
        LAUNCH(h, k_collect_energies, cdiv(h->nE, 128), 128, 0, h->nE, kESlots, h->out.p + n3, d_eq.p + (size_t)h->nE * b);
    }
    std::vector<double> eq((size_t)h->nE * nbeads);
    CU(cudaMemcpyAsync(eq.data(), d_eq.p, sizeof(double) * eq.size(), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    for (int b = 0; b < nbeads; b++)
        for (int k = 0; k < nEQ; k++) EQ_out[(size_t)b * nEQ + k] = eq[(size_t)b * h->nE + QNB_E_COUNT + k];
    return 0;
}

int qnb_build_lists(qnb_handle *h, const double *x, double Rq, double Rcq2, double RcLRF2, double Rcpp2, double Rcpw2,
                    double Rcww2, double RcLRF, int64_t counts_out[8]) {
    if (!h || !x) return fail("qnb_build_lists: null argument");
    CU(cudaSetDevice(h->device));
    const qnb_system &s = h->T.s;
    if (s.use_PBC && !(h->box[0] > 0 && h->box[1] > 0 && h->box[2] > 0)) return fail("qnb_build_lists: periodic system without a box (call qnb_update_box)");
    Cut &C = h->cut;
    C.rc2[0] = Rcpp2; C.rc2[1] = Rcpw2; C.rc2[2] = Rcww2;
    C.rclrf2 = RcLRF2; C.rcq2 = Rcq2; C.Rq = Rq;
    // "no LRF cut-off" sentinels of the box builders: pp compares its squared argument with -1
    // (nonbondene.f90:2035), pw/ww compare the global RcLRF (L3066, L4497)
    C.lrf_all[0] = s.use_PBC && (RcLRF2 == -1.0);
    C.lrf_all[1] = C.lrf_all[2] = s.use_PBC && (RcLRF == -1.0);
    const size_t n3 = 3 * (size_t)s.natom;
    memcpy(h->hx, x, n3 * sizeof(double));
    CU(cudaMemcpyAsync(h->x.p, h->hx, n3 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    h->x_from_nonbond = false;
    if (build_device(h, x)) return 1;
    if (counts_out) {
        for (int k = 0; k < 8; k++) counts_out[k] = 0;
        int64_t n;
        for (int w = 0; w < 5; w++) { if (qnb_list_count(h, w, 1, &n)) return 1; counts_out[w] = n; }
    }
    return 0;
}

// make_pair_lists of several independent systems: every build has host round trips for its sizes, so each runs on its
// own host thread (the handles own their streams; graph capture is thread-local) and the builds overlap on the device.
int qnb_build_lists_batch(int n, qnb_handle *const *hs, const double *const *x, double Rq, double Rcq2, double RcLRF2, double Rcpp2,
                          double Rcpw2, double Rcww2, double RcLRF, int64_t *counts_out /* [n][8] or NULL */) {
    if (n < 0 || (n > 0 && (!hs || !x))) return fail("qnb_build_lists_batch: null argument");
    for (int k = 0; k < n; k++) {
        if (!hs[k] || !x[k]) return fail("qnb_build_lists_batch: null argument for system %d", k);
        for (int j = 0; j < k; j++) if (hs[j] == hs[k]) return fail("qnb_build_lists_batch: handle %d given twice", k);
    }
    static const int share_cap = [] { const char *e = getenv("QNB_BATCH_SHARE"); return e ? std::max(1, atoi(e)) : 8; }();
    for (int k = 0; k < n; k++) hs[k]->share = std::min(n, share_cap);
    std::vector<int> rc(n, 0);
    std::vector<std::string> err(n);
    auto work = [&](int k) {
        rc[k] = qnb_build_lists(hs[k], x[k], Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF, counts_out ? counts_out + 8 * (size_t)k : nullptr);
        if (rc[k]) err[k] = g_err;
    };
    std::vector<std::thread> th;
    for (int k = 1; k < n; k++) th.emplace_back(work, k);
    if (n > 0) work(0);
    for (auto &t : th) t.join();
    for (int k = 0; k < n; k++) if (rc[k]) return fail("qnb_build_lists_batch: system %d: %s", k, err[k].c_str());
    return 0;
}

// qnb_nonbond in two halves, so that several handles can be in flight at once (qnb_nonbond_batch)
static int nonbond_begin(qnb_handle *h, const double *x, const double *lambda, int flags, double *d) {
    CU(cudaSetDevice(h->device));
    const qnb_system &s = h->T.s;
    const size_t n3 = 3 * (size_t)s.natom;
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    h->x_direct = host_alias(h, x, n3 * sizeof(double)) ? x : nullptr;
    h->d_direct = static_cast<double *>(host_alias(h, d, n3 * sizeof(double)));
    h->d_host = d;
    h->d_overwrite = h->d_direct && (flags & QNB_FLAG_D_IS_ZERO);
    if (!h->x_direct) memcpy(h->hx, x, n3 * sizeof(double));
    for (int k = 0; k < s.nstates; k++) h->hlam[k] = lambda[k];
    h->last_flags = flags;
    const auto t1 = clk::now();
    const int rc = step_device(h, flags, true);
    h->t_stage_in = std::chrono::duration<double>(t1 - t0).count();
    h->t_issue = std::chrono::duration<double>(clk::now() - t1).count();
    return rc;
}
static int nonbond_end(qnb_handle *h, double *d, double *E_out, double *EQ_out) {
    const qnb_system &s = h->T.s;
    const size_t n3 = 3 * (size_t)s.natom;
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    h->last_h2d = (int64_t)((n3 + s.nstates) * sizeof(double));
    h->last_d2h = (int64_t)(h->nout * sizeof(double));
    if (!h->d_direct) {
        const double *__restrict__ g = h->hout;
        for (size_t k = 0; k < n3; k++) d[k] += g[k];
    }
    for (int k = 0; k < h->nE; k++) {
        double e = 0;   // fixed-order sum of the partial accumulators
        for (int sl = 0; sl < kESlots; sl++) e += h->hout[n3 + (size_t)sl * h->nE + k];
        if (k < QNB_E_COUNT) E_out[k] = e; else EQ_out[k - QNB_E_COUNT] = e;
    }
    for (int k = 0; k < kRstOut; k++) h->last_rst[k] = h->hout[n3 + (size_t)kESlots * h->nE + k];
    h->x_from_nonbond = true;
    return 0;
}

int qnb_nonbond(qnb_handle *h, const double *x, const double *lambda, int flags, double *d, double *E_out, double *EQ_out) {
    if (!h || !x || !lambda || !d || !E_out || !EQ_out) return fail("qnb_nonbond: null argument");
    if (!h->lists_built) return fail("qnb_nonbond: pair lists have not been built (call qnb_build_lists)");
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    if (nonbond_begin(h, x, lambda, flags, d)) return 1;
    const auto t2 = clk::now();
    CU(cudaStreamSynchronize(h->st));
    const auto t3 = clk::now();
    if (nonbond_end(h, d, E_out, EQ_out)) return 1;
    const auto t4 = clk::now();
    auto sec = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    (void)t0;
    h->t_wait = sec(t2, t3); h->t_add_out = sec(t3, t4);
    return 0;
}

// Independent systems of one process (FEP lambda windows, EVB frames, replicas) advanced together: the step of every
// handle is in flight on its own streams before the first result is awaited, so the kernels of the windows fill the SMs
// side by side and the staging of window k+1 overlaps the device work of window k.  Same results as n qnb_nonbond calls.
// Host-side workers of qnb_nonbond_batch: staging x into pinned memory, launching the graph and adding the gradient out
// are ~25 us of host work per window (C2), serial in the caller's thread they would bound a batch of 7 at ~180 us.  A
// small persistent pool (created on first use) shares the windows; the calling thread works too.
namespace {
struct BatchPool {
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::vector<std::thread> th;
    std::function<void(int)> job;   // job(worker index 1..nworker)
    unsigned long long gen = 0;
    int pending = 0;
    bool stop = false;
    ~BatchPool() {
        { std::lock_guard<std::mutex> l(mu); stop = true; gen++; }
        cv_go.notify_all();
        for (auto &t : th) t.join();
    }
    void ensure(int nworker) {
        while ((int)th.size() < nworker) {
            const int id = (int)th.size() + 1;
            const unsigned long long start_gen = gen;
            th.emplace_back([this, id, start_gen] {
                unsigned long long seen = start_gen;
                for (;;) {
                    std::function<void(int)> f;
                    {
                        std::unique_lock<std::mutex> l(mu);
                        cv_go.wait(l, [&] { return gen != seen; });
                        seen = gen;
                        if (stop) return;
                        f = job;
                    }
                    if (f) f(id);
                    { std::lock_guard<std::mutex> l(mu); if (--pending == 0) cv_done.notify_one(); }
                }
            });
        }
    }
    // run f(0) here and f(1..nworker) on the pool; returns when all are done
    void run(int nworker, const std::function<void(int)> &f) {
        ensure(nworker);
        { std::lock_guard<std::mutex> l(mu); job = [f, nworker](int id) { if (id <= nworker) f(id); }; pending = (int)th.size(); gen++; }
        cv_go.notify_all();
        f(0);
        std::unique_lock<std::mutex> l(mu);
        cv_done.wait(l, [&] { return pending == 0; });
    }
};
BatchPool g_pool;
std::mutex g_pool_use;
double g_batch_t[3] = {0, 0, 0};   // calling thread, last batch: issuing (staging + launches), waiting for the device, adding out
}  // namespace

int qnb_bench_last_batch_timing(double out[3]) {
    if (!out) return fail("null argument");
    for (int k = 0; k < 3; k++) out[k] = g_batch_t[k];
    return 0;
}

// ---- the batched step as a parent graph whose children are the windows' own step graphs
namespace {
struct ParentPlan {
    std::vector<qnb_handle *> hs;
    std::vector<uint64_t> serial;
    std::vector<const double *> xptr;
    std::vector<double *> dptr;
    std::vector<char> dow;
    int flags = -1, nodes = 0;
    cudaGraphExec_t ge = nullptr;
    bool valid = false;
};
ParentPlan g_parent;
}  // namespace

namespace qnb {
// returns 0 / 1 like qnb_nonbond_batch, or -1 when the batch has to take the per-window path
static int batch_parent_run(int n, qnb_handle *const *hs, const double *const *x, const double *const *lambda, int flags,
                            double *const *d, double *const *E_out, double *const *EQ_out) {
    ParentPlan &B = g_parent;
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    if (cudaSetDevice(hs[0]->device) != cudaSuccess) return -1;
    const int kflags = flags & (QNB_FLAG_MD | QNB_FLAG_QQ | QNB_FLAG_NO_ENERGY | QNB_FLAG_SOLVENT_RESTRAINTS);
    // stage what nonbond_begin stages, without launching
    for (int w = 0; w < n; w++) {
        qnb_handle *h = hs[w];
        const size_t n3 = 3 * (size_t)h->T.s.natom;
        h->x_direct = host_alias(h, x[w], n3 * sizeof(double)) ? x[w] : nullptr;
        h->d_direct = static_cast<double *>(host_alias(h, d[w], n3 * sizeof(double)));
        h->d_host = d[w];
        h->d_overwrite = h->d_direct && (flags & QNB_FLAG_D_IS_ZERO);
        if (!h->x_direct) memcpy(h->hx, x[w], n3 * sizeof(double));
        for (int k = 0; k < h->T.s.nstates; k++) h->hlam[k] = lambda[w][k];
        h->last_flags = flags;
    }
    bool same = B.valid && (int)B.hs.size() == n && B.flags == kflags;
    for (int w = 0; same && w < n; w++)
        same = B.hs[w] == hs[w] && B.serial[w] == hs[w]->build_serial && B.xptr[w] == hs[w]->x_direct && B.dptr[w] == hs[w]->d_direct &&
               (bool)B.dow[w] == hs[w]->d_overwrite;
    if (!same) {
        B.valid = false;
        cudaGraph_t parent = nullptr;
        if (cudaGraphCreate(&parent, 0) != cudaSuccess) { cudaGetLastError(); return -1; }
        int nodes = 0;
        bool ok = true;
        std::vector<cudaGraph_t> kids;
        for (int w = 0; w < n && ok; w++) {
            cudaGraph_t g = nullptr;
            int nn = 0;
            if (step_device(hs[w], kflags, true, &g, &nn) || !g) { ok = false; break; }
            kids.push_back(g);
            cudaGraphNode_t node;
            ok = cudaGraphAddChildGraphNode(&node, parent, nullptr, 0, g) == cudaSuccess;   // no dependencies: the windows run side by side
            nodes += nn;
        }
        if (ok) {
            bool updated = false;
            if (B.ge && (int)B.hs.size() == n) {
                cudaGraphExecUpdateResultInfo info;
                updated = cudaGraphExecUpdate(B.ge, parent, &info) == cudaSuccess;
                if (!updated) cudaGetLastError();
            }
            if (!updated) {
                if (B.ge) { cudaGraphExecDestroy(B.ge); B.ge = nullptr; }
                ok = cudaGraphInstantiate(&B.ge, parent, 0) == cudaSuccess;
            }
        }
        for (cudaGraph_t g : kids) cudaGraphDestroy(g);
        cudaGraphDestroy(parent);
        if (!ok) { cudaGetLastError(); if (B.ge) { cudaGraphExecDestroy(B.ge); B.ge = nullptr; } B.hs.clear(); return -1; }
        B.hs.assign(hs, hs + n); B.serial.resize(n); B.xptr.resize(n); B.dptr.resize(n); B.dow.resize(n);
        for (int w = 0; w < n; w++) { B.serial[w] = hs[w]->build_serial; B.xptr[w] = hs[w]->x_direct; B.dptr[w] = hs[w]->d_direct; B.dow[w] = hs[w]->d_overwrite; }
        B.flags = kflags; B.nodes = nodes; B.valid = true;
    }
    CU(cudaGraphLaunch(B.ge, hs[0]->st));
    hs[0]->launches += B.nodes;
    const auto t1 = clk::now();
    CU(cudaStreamSynchronize(hs[0]->st));
    const auto t2 = clk::now();
    int rc = 0;
    for (int w = 0; w < n; w++)
        if (nonbond_end(hs[w], d[w], E_out[w], EQ_out[w])) rc = 1;
    g_batch_t[0] = std::chrono::duration<double>(t1 - t0).count();
    g_batch_t[1] = std::chrono::duration<double>(t2 - t1).count();
    g_batch_t[2] = std::chrono::duration<double>(clk::now() - t2).count();
    return rc;
}
}  // namespace qnb

// ---- the batched step as ONE graph: every kernel type launched once for all windows (k_batched, blockIdx.z = window)
namespace {
constexpr int kSlotPack = qnb::K_COUNT, kSlotCollect = qnb::K_COUNT + 1, kSlots = qnb::K_COUNT + 2;
struct BatchPlan {
    std::vector<qnb_handle *> hs;
    std::vector<uint64_t> serial;
    std::vector<const double *> xptr;
    std::vector<double *> dptr;
    int flags = -1, nEmax = 0, nodes = 0;
    qnb::BatchSlot slots[kSlots];
    bool active[qnb::K_COUNT] = {};
    cudaGraphExec_t ge = nullptr;
    double *hlam = nullptr, *hE = nullptr;   // pinned: lambda of all windows in, [E | EQ] of all windows out
    qnb::DBuf<double> dlam, dE;
    size_t cap = 0;
    bool valid = false;
};
BatchPlan g_plan;   // the last batch (a FEP farm advances the same windows step after step)
std::mutex g_plan_mu;
}  // namespace

namespace qnb {
// true if the batch can run as one graph; fills / refreshes g_plan
static bool batch_plan_ready(int n, qnb_handle *const *hs, const double *const *x, int flags, double *const *d) {
    // Opt-in (QNB_BATCH_GRAPH=1).  Measured r02w-r02y, 7 windows of C2: 351-375 us per batched step against 236 us for
    // seven per-window graphs on their own streams.  One launch per kernel type does halve the per-window time of every
    // kernel (water rows 12.3 -> 6.3 us, Q partner 13 -> 4.8, Q atom 16.5 -> 8.5, solute rows 12.4 -> 7.1), but in one graph
    // the seven uploads and downloads (~120 us) no longer overlap the other windows' kernels, and the argument tuples read
    // from shared memory cost registers (pair energies 56 -> 156 without a cap, spills with one).
    const bool on = getenv("QNB_BATCH_GRAPH") != nullptr;   // read every call: tests switch it
    if (!on || n < 2 || !(flags & QNB_FLAG_D_IS_ZERO) || !(flags & QNB_FLAG_MD) || (flags & QNB_FLAG_SOLVENT_RESTRAINTS)) return false;
    BatchPlan &B = g_plan;
    const int kflags = flags & (QNB_FLAG_MD | QNB_FLAG_QQ | QNB_FLAG_NO_ENERGY);
    for (int w = 0; w < n; w++) {
        qnb_handle *h = hs[w];
        if (h->device != hs[0]->device || !h->use_graph || !h->multi_stream || h->comm || p2p_ready(h) || h->legacy_rows || h->npk <= 0) return false;
        const size_t bytes = 3 * (size_t)h->T.s.natom * sizeof(double);
        if (!host_alias(h, x[w], bytes) || !host_alias(h, d[w], bytes)) return false;
    }
    bool same = B.valid && (int)B.hs.size() == n && B.flags == kflags;
    for (int w = 0; same && w < n; w++)
        same = B.hs[w] == hs[w] && B.serial[w] == hs[w]->build_serial && B.xptr[w] == x[w] && B.dptr[w] == d[w];
    if (same) return true;
    // ---- (re)build: record every window's launch arguments through the ordinary launch sites
    B.valid = false;
    B.hs.assign(hs, hs + n); B.serial.resize(n); B.xptr.assign(x, x + n); B.dptr.assign(d, d + n); B.flags = kflags;
    for (auto &sl : B.slots) sl.reset();
    B.nEmax = 0;
    for (int w = 0; w < n; w++) B.nEmax = std::max(B.nEmax, hs[w]->nE);
    if (cudaSetDevice(hs[0]->device) != cudaSuccess) return false;
    if ((size_t)n > B.cap) {
        if (B.hlam) cudaFreeHost(B.hlam);
        if (B.hE) cudaFreeHost(B.hE);
        B.hlam = B.hE = nullptr;
        if (cudaMallocHost(&B.hlam, (size_t)n * kMaxStates * sizeof(double)) != cudaSuccess) return false;
        if (cudaMallocHost(&B.hE, (size_t)n * 256 * sizeof(double)) != cudaSuccess) return false;
        B.cap = n;
    }
    if (B.nEmax > 256 || B.dlam.ensure((size_t)n * kMaxStates) || B.dE.ensure((size_t)n * 256)) return false;
    for (int k = 0; k < K_COUNT; k++) B.active[k] = step_kernel_active(hs[0], k, kflags);
    for (int w = 0; w < n; w++) {
        qnb_handle *h = hs[w];
        B.serial[w] = h->build_serial;
        for (int k = 0; k < K_COUNT; k++)
            if (step_kernel_active(h, k, kflags) != B.active[k]) return false;
        collect_launch<PackStepBody, 256, 1>(B.slots[kSlotPack], dim3(cdiv(h->npk, 256)), 0, h->npk, h->fix, h->D.nat_solute, (const int *)h->pk_atom.p,
                                             (const int *)h->pk_sw.p, (const float *)h->pk_q.p, (const int *)h->pk_ct.p, (const double *)h->x.p, h->px.p, h->py.p,
                                             h->pz.p, h->rec_i.p, h->rec_f.p, h->wT.p, h->wown.p, h->wd.p, h->out.p, (int)h->nout,
                                             (const double *)(B.dlam.p + (size_t)w * kMaxStates), h->lam_dev, (int)kMaxStates);
        for (int k = 0; k < K_COUNT; k++) {
            if (!B.active[k]) continue;
            h->collect = &B.slots[k];
            launch_step_kernel(h, k, nullptr, kflags);
            h->collect = nullptr;
        }
        collect_launch<CollectEnergiesBody, 128, 1>(B.slots[kSlotCollect], dim3(cdiv(h->nE, 128)), 0, h->nE, (int)kESlots,
                                                    (const double *)(h->out.p + 3 * (size_t)h->T.s.natom), B.dE.p + (size_t)w * 256);
    }
    for (int k = 0; k < kSlots; k++) {
        BatchSlot &sl = B.slots[k];
        if (sl.bad) return false;
        if (!sl.launch) continue;
        if ((int)sl.grids.size() != n) return false;
        if (sl.dargs.ensure(sl.args.size()) || sl.dgrids.ensure(sl.grids.size())) return false;
        if (cudaMemcpy(sl.dargs.p, sl.args.data(), sl.args.size(), cudaMemcpyHostToDevice) != cudaSuccess) return false;
        if (cudaMemcpy(sl.dgrids.p, sl.grids.data(), sl.grids.size() * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    }
    // ---- capture: uploads, pack (+ clearing, + lambda), the step kernels side by side, energy sums, downloads
    qnb_handle *h0 = hs[0];
    cudaStream_t st = h0->st;
    cudaGraph_t g = nullptr;
    int rc = 0, nodes = 0;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    for (int w = 0; w < n; w++) {
        rc |= cudaMemcpyAsync(hs[w]->x.p, x[w], 3 * (size_t)hs[w]->T.s.natom * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess;
        nodes++;
    }
    rc |= cudaMemcpyAsync(B.dlam.p, B.hlam, (size_t)n * kMaxStates * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess;
    auto fire = [&](int k, cudaStream_t cs) {
        BatchSlot &sl = B.slots[k];
        sl.launch(sl.dargs.p, sl.dgrids.p, dim3(sl.grid.x, sl.grid.y, (unsigned)n), sl.smem, cs);
        nodes++;
    };
    fire(kSlotPack, st);
    rc |= cudaEventRecord(h0->ev_fork, st) != cudaSuccess;
    bool used[kAux] = {};
    static const int kOrder[K_COUNT] = {K_RST, K_QSTATIC, K_SOLUTE, K_QATOM, K_QPARTNER, K_WATER, K_ENERGY, K_LRF};
    for (int o = 0; o < K_COUNT; o++) {
        const int k = kOrder[o];
        if (!B.active[k] || !B.slots[k].launch) continue;
        const int si = kStreamOf[k];
        cudaStream_t cs = si < 0 ? st : h0->aux[si];
        if (si >= 0 && !used[si]) { rc |= cudaStreamWaitEvent(cs, h0->ev_fork, 0) != cudaSuccess; used[si] = true; }
        fire(k, cs);
    }
    for (int k = 0; k < kAux; k++)
        if (used[k]) {
            rc |= cudaEventRecord(h0->ev_join[k], h0->aux[k]) != cudaSuccess;
            rc |= cudaStreamWaitEvent(st, h0->ev_join[k], 0) != cudaSuccess;
        }
    fire(kSlotCollect, st);
    for (int w = 0; w < n; w++) {
        rc |= cudaMemcpyAsync(d[w], hs[w]->out.p, 3 * (size_t)hs[w]->T.s.natom * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess;
        nodes++;
    }
    rc |= cudaMemcpyAsync(B.hE, B.dE.p, (size_t)n * 256 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess;
    nodes += 2;
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    if (rc || ce != cudaSuccess || !g || cudaGetLastError() != cudaSuccess) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return false; }
    bool updated = false;
    if (B.ge) {
        cudaGraphExecUpdateResultInfo info;
        updated = cudaGraphExecUpdate(B.ge, g, &info) == cudaSuccess;
        if (!updated) { cudaGetLastError(); cudaGraphExecDestroy(B.ge); B.ge = nullptr; }
    }
    if (!updated) ce = cudaGraphInstantiate(&B.ge, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) { cudaGetLastError(); B.ge = nullptr; return false; }
    B.nodes = nodes;
    B.valid = true;
    return true;
}

static int batch_plan_run(int n, qnb_handle *const *hs, const double *const *lambda, int flags, double *const *E_out, double *const *EQ_out) {
    BatchPlan &B = g_plan;
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    for (int w = 0; w < n; w++) {
        for (int k = 0; k < kMaxStates; k++) B.hlam[(size_t)w * kMaxStates + k] = k < hs[w]->T.s.nstates ? lambda[w][k] : 0.0;
        hs[w]->last_flags = flags;
    }
    CU(cudaSetDevice(hs[0]->device));
    CU(cudaGraphLaunch(B.ge, hs[0]->st));
    const auto t1 = clk::now();
    CU(cudaStreamSynchronize(hs[0]->st));
    CU(cudaGetLastError());
    const auto t2 = clk::now();
    for (int w = 0; w < n; w++) {
        qnb_handle *h = hs[w];
        const double *e = B.hE + (size_t)w * 256;
        for (int k = 0; k < h->nE; k++) { if (k < QNB_E_COUNT) E_out[w][k] = e[k]; else EQ_out[w][k - QNB_E_COUNT] = e[k]; }
        const size_t n3 = 3 * (size_t)h->T.s.natom;
        h->last_h2d = (int64_t)((n3 + h->T.s.nstates) * sizeof(double));
        h->last_d2h = (int64_t)((n3 + h->nE) * sizeof(double));
        h->x_from_nonbond = true;
    }
    hs[0]->launches += B.nodes;
    g_batch_t[0] = std::chrono::duration<double>(t1 - t0).count();
    g_batch_t[1] = std::chrono::duration<double>(t2 - t1).count();
    g_batch_t[2] = std::chrono::duration<double>(clk::now() - t2).count();
    return 0;
}
}  // namespace qnb

int qnb_nonbond_batch(int n, qnb_handle *const *hs, const double *const *x, const double *const *lambda, int flags,
                      double *const *d, double *const *E_out, double *const *EQ_out) {
    if (n < 0 || (n > 0 && (!hs || !x || !lambda || !d || !E_out || !EQ_out))) return fail("qnb_nonbond_batch: null argument");
    for (int k = 0; k < n; k++) {
        if (!hs[k] || !x[k] || !lambda[k] || !d[k] || !E_out[k] || !EQ_out[k]) return fail("qnb_nonbond_batch: null argument for system %d", k);
        if (!hs[k]->lists_built) return fail("qnb_nonbond_batch: pair lists of system %d have not been built", k);
        for (int j = 0; j < k; j++) if (hs[j] == hs[k]) return fail("qnb_nonbond_batch: handle %d given twice", k);
    }
    // measured (r02l, C2 x 7): 4 host threads made the batched step SLOWER (0.74 vs 0.52 ms; the CUDA runtime serialises
    // the launches and the wake-ups cost more than the copies they spread), so one thread is the default
    // every kernel type once for all windows, one graph launch (needs registered x and d and QNB_FLAG_D_IS_ZERO)
    {
        std::lock_guard<std::mutex> lock(g_plan_mu);
        if (qnb::batch_plan_ready(n, hs, x, flags, d)) return qnb::batch_plan_run(n, hs, lambda, flags, E_out, EQ_out);
    }
    // ---- opt-in (QNB_BATCH_PARENT=1): the windows' step graphs as children of ONE parent graph, one launch instead of n.
    // Measured r03o (7 windows): no gain -- launching the parent costs what its ~85 nodes cost (49 us against 50-58 us for
    // seven launches), C3 176 vs 167 us per batched step, C2 241 vs 236.
    const bool use_parent = getenv("QNB_BATCH_PARENT") != nullptr;
    if (use_parent && n >= 2) {
        bool ok = true;
        for (int k = 0; k < n && ok; k++) ok = hs[k]->use_graph && !hs[k]->comm && hs[k]->device == hs[0]->device;
        if (ok) {
            std::lock_guard<std::mutex> lock(g_plan_mu);
            const int rc = qnb::batch_parent_run(n, hs, x, lambda, flags, d, E_out, EQ_out);
            if (rc >= 0) return rc;   // -1: not applicable, fall through to the per-window launches
        }
    }
    static const int max_workers = [] { const char *e = getenv("QNB_BATCH_THREADS"); return e ? std::max(1, atoi(e)) : 1; }();
    const int nw = std::max(1, std::min(max_workers, n));   // threads incl. the caller
    std::vector<int> rc(nw, 0);
    std::vector<std::string> err(nw);
    // thread w takes windows w, w + nw, ...: all of its steps are in flight before it waits for the first result
    using clk = std::chrono::steady_clock;
    auto work = [&](int w) {
        int started = 0;
        const auto t0 = clk::now();
        for (int k = w; k < n; k += nw, started++)
            if (nonbond_begin(hs[k], x[k], lambda[k], flags, d[k])) { rc[w] = 1; err[w] = g_err; break; }
        const auto t1 = clk::now();
        double wait = 0.0;
        for (int k = w, j = 0; j < started; k += nw, j++) {
            const auto a = clk::now();
            cudaStreamSynchronize(hs[k]->st);
            wait += std::chrono::duration<double>(clk::now() - a).count();
            if (nonbond_end(hs[k], d[k], E_out[k], EQ_out[k]) && !rc[w]) { rc[w] = 1; err[w] = g_err; }
        }
        if (w == 0) {
            g_batch_t[0] = std::chrono::duration<double>(t1 - t0).count();
            g_batch_t[1] = wait;
            g_batch_t[2] = std::chrono::duration<double>(clk::now() - t1).count() - wait;
        }
    };
    if (nw == 1) work(0);
    else {
        std::lock_guard<std::mutex> use(g_pool_use);
        g_pool.run(nw - 1, work);
    }
    for (int w = 0; w < nw; w++) if (rc[w]) return fail("%s", err[w].c_str());
    return 0;
}

int qnb_register_host_buffers(qnb_handle *h, const double *x, double *d) {
    if (!h) return fail("null handle");
    CU(cudaSetDevice(h->device));
    const size_t bytes = 3 * (size_t)h->T.s.natom * sizeof(double);
    if (x && !host_alias(h, x, bytes, true) && h->host_direct) return fail("qnb_register_host_buffers: cudaHostRegister(x) failed");
    if (d && !host_alias(h, d, bytes, true) && h->host_direct) return fail("qnb_register_host_buffers: cudaHostRegister(d) failed");
    return 0;
}

int qnb_release_host_buffers(qnb_handle *h) {
    if (!h) return fail("null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->st));
    release_host_buffers(h);
    return 0;
}

int qnb_last_timing(qnb_handle *h, double out[4]) {
    if (!h || !out) return fail("null argument");
    out[0] = h->t_stage_in; out[1] = h->t_issue; out[2] = h->t_wait; out[3] = h->t_add_out;
    return 0;
}

// ---- list export: expansion of the device rows into the reference's explicit entries (host side)
static int export_impl(qnb_handle *h, int which, int state, int32_t *ij, double *params, int64_t capacity, int64_t *count) {
    const HostTables &T = h->T;
    const qnb_system &s = T.s;
    const int nst = s.nstates;
    int64_t n = 0;
    auto emit = [&](int i1, int j1, const QPar &p) -> bool {
        if (ij) {
            if (n >= capacity) return false;
            ij[2 * n] = i1; ij[2 * n + 1] = j1;
            if (params) { params[4 * n] = p.A; params[4 * n + 1] = p.B; params[4 * n + 2] = p.el; params[4 * n + 3] = p.score; }
        }
        n++;
        return true;
    };
    if (state < 1 || state > nst) return fail("state %d out of range", state);
    if (which == QNB_LIST_QQ || which == QNB_LIST_QQP) {
        if (s.is_master)
            for (const auto &e : (which == QNB_LIST_QQ ? T.qq_list : T.qqp_list))
                if (e.state == state - 1)
                    if (!emit(e.iq, which == QNB_LIST_QQ ? e.jq : e.j + 1, e.p)) return fail("export capacity too small");
        *count = n;
        return 0;
    }
    if (!h->lists_built) return fail("pair lists have not been built");
    CU(cudaSetDevice(h->device));
    if (which == QNB_LIST_QP || which == QNB_LIST_QW) {
        if (s.nqat > 0) {
            const int m = which == QNB_LIST_QP ? h->nqp : h->nqw;
            std::vector<int> lst(std::max(m, 1));
            if (m) CU(cudaMemcpy(lst.data(), which == QNB_LIST_QP ? h->qp_list.p : h->qw_list.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
            for (int k = 0; k < m; k++)
                for (int q = 1; q <= s.nqat; q++) {
                    if (which == QNB_LIST_QP) {
                        const QPar &p = T.qp_tab[((size_t)(q - 1) * nst + (state - 1)) * s.nat_solute + lst[k]];
                        if (!emit(q, lst[k] + 1, p)) return fail("export capacity too small");
                    } else
                        for (int site = 0; site < s.solv_atom; site++) {
                            const QPar &p = T.qw_tab[((size_t)(q - 1) * nst + (state - 1)) * s.solv_atom + site];
                            if (!emit(q, s.nat_solute + s.solv_atom * lst[k] + site + 1, p)) return fail("export capacity too small");
                        }
                }
        }
        *count = n;
        return 0;
    }
    const int nu = T.nunit;
    std::vector<int> counts(3 * (size_t)std::max(nu, 1)), off(std::max(nu, 1) + 1);
    std::vector<uint32_t> rows((size_t)std::max<int64_t>(h->total_rows, 1));
    if (nu) {
        CU(cudaMemcpy(counts.data(), h->counts.p, sizeof(int) * 3 * nu, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(off.data(), h->row_off.p, sizeof(int) * (nu + 1), cudaMemcpyDeviceToHost));
        if (h->total_rows) CU(cudaMemcpy(rows.data(), h->rows.p, sizeof(uint32_t) * h->total_rows, cudaMemcpyDeviceToHost));
    }
    const int ns = s.ncgp_solute, sa = s.solv_atom;
    std::vector<int> pk(std::max(h->npk, 1));   // row entries are packed atom indices
    if (h->npk) CU(cudaMemcpy(pk.data(), h->pk_atom.p, sizeof(int) * h->npk, cudaMemcpyDeviceToHost));
    if (which == QNB_LIST_WW) {
        for (int u = ns; u < nu; u++) {
            const int iw = u - ns;
            for (int k = 0; k < counts[3 * u]; k++) {
                const int jw = (pk[rows[off[u] + k] & kIdMask] - s.nat_solute) / sa;
                for (int la = 0; la < sa; la++)
                    for (int ka = 0; ka < sa; ka++)
                        if (!emit(s.nat_solute + sa * iw + la + 1, s.nat_solute + sa * jw + ka + 1, T.ww_par[la * sa + ka]))
                            return fail("export capacity too small");
            }
        }
    } else if (which == QNB_LIST_PW) {
        for (int g = 0; g < ns; g++) {
            const uint32_t *rb = rows.data() + off[g] + counts[3 * g] + counts[3 * g + 1];
            for (int k = 0; k < counts[3 * g + 2]; k++) {
                const int jw = (pk[rb[k] & kIdMask] - s.nat_solute) / sa;
                for (int m = 0; m < T.g_n[g]; m++) {
                    const int i = T.g_atoms[T.g_first[g] + m];
                    if (T.is_q[i]) continue;
                    for (int site = 0; site < sa; site++)
                        if (!emit(i + 1, s.nat_solute + sa * jw + site + 1, T.pw_params(i, site))) return fail("export capacity too small");
                }
            }
        }
    } else if (which == QNB_LIST_PP) {
        for (int g = 0; g < ns; g++) {
            const uint32_t *ra = rows.data() + off[g];
            for (int m = 0; m < T.g_n[g]; m++) {
                const int i = T.g_atoms[T.g_first[g] + m];
                if (T.is_q[i]) continue;
                for (int k = 0; k < counts[3 * g]; k++) {
                    const int j = pk[ra[k] & kIdMask];
                    if (T.grp_of_atom[j] == g && i >= j) continue;   // count once inside a group (L1874)
                    QPar p;
                    bool set;
                    if (!T.pp_params(i, j, p, set) || !set) continue;   // pp_map == 0 or .not. %set (L1876-1877)
                    if (!emit(i + 1, j + 1, p)) return fail("export capacity too small");
                }
            }
        }
    } else return fail("unknown list %d", which);
    *count = n;
    return 0;
}

int qnb_list_count(qnb_handle *h, int which, int state, int64_t *n) {
    if (!h || !n) return fail("null argument");
    return export_impl(h, which, state, nullptr, nullptr, 0, n);
}

int qnb_export_list(qnb_handle *h, int which, int state, int32_t *ij, double *params, int64_t capacity) {
    if (!h || !ij) return fail("null argument");
    int64_t n;
    return export_impl(h, which, state, ij, params, capacity, &n);
}

int qnb_export_lrf(qnb_handle *h, double *lrf) {
    if (!h || !lrf) return fail("null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpy(lrf, h->lrf.p, sizeof(double) * QNB_LRF_STRIDE * h->T.s.ncgp, cudaMemcpyDeviceToHost));
    return 0;
}

int qnb_comm_unique_id(void *id128) {
    if (!g_nccl.load()) return fail("libnccl.so.2 not found: %s", dlerror());
    ncclUniqueIdT id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc) return fail("ncclGetUniqueId failed (%d)", rc);
    memcpy(id128, &id, 128);
    return 0;
}

int qnb_comm_init(qnb_handle *h, int rank, int nranks, const void *id128) {
    if (!h || !id128) return fail("null argument");
    if (!g_nccl.load()) return fail("libnccl.so.2 not found: %s", dlerror());
    CU(cudaSetDevice(h->device));
    ncclUniqueIdT id;
    memcpy(&id, id128, 128);
    int rc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (rc) return fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    h->rank = rank; h->nranks = nranks;
    // NCCL takes over on EVERY rank: a peer-memory attach that succeeded here but failed on another rank (the host falls
    // back to this call on all of them) must not leave this rank on the peer-memory kernel
    if (h->p2p.n > 1) {
        for (int k = 0; k < kMaxPeers; k++)
            if (h->peer_base[k]) { cudaIpcCloseMemHandle(h->peer_base[k]); h->peer_base[k] = nullptr; }
        h->p2p = P2PComm{};
        drop_graphs(h);
    }
    return 0;
}

// ---- all-reduce over peer memory (NVLink / NVSwitch): the ranks of one node map each other's arenas through CUDA IPC
struct IpcBlob { cudaIpcMemHandle_t mem; int32_t device; int32_t pad; uint64_t doubles, lrf_off, ctl_off; };
static_assert(sizeof(IpcBlob) <= QNB_IPC_BLOB, "IPC blob");

int qnb_comm_ipc_export(qnb_handle *h, void *blob) {
    if (!h || !blob) return fail("null argument");
    CU(cudaSetDevice(h->device));
    IpcBlob b{};
    CU(cudaIpcGetMemHandle(&b.mem, h->arena));
    b.device = h->device; b.doubles = h->arena_doubles; b.lrf_off = h->arena_lrf_off; b.ctl_off = h->arena_ctl_off;
    memset(blob, 0, QNB_IPC_BLOB);
    memcpy(blob, &b, sizeof b);
    return 0;
}

int qnb_comm_ipc_attach(qnb_handle *h, int rank, int nranks, const void *blobs) {
    if (!h || !blobs) return fail("null argument");
    if (nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks) return fail("qnb_comm_ipc_attach: rank %d of %d (at most %d ranks)", rank, nranks, kMaxPeers);
    CU(cudaSetDevice(h->device));
    P2PComm c{};
    c.rank = rank; c.n = nranks; c.ctl_off = h->arena_ctl_off;
    for (int p = 0; p < nranks; p++) {
        IpcBlob b;
        memcpy(&b, (const char *)blobs + (size_t)p * QNB_IPC_BLOB, sizeof b);
        if (b.doubles != h->arena_doubles || b.lrf_off != h->arena_lrf_off || b.ctl_off != h->arena_ctl_off)
            return fail("qnb_comm_ipc_attach: rank %d holds a different system (arena layout differs)", p);
        if (p == rank) { c.base[p] = h->arena; continue; }
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, h->device, b.device));
        if (!can) return fail("qnb_comm_ipc_attach: device %d cannot access device %d of rank %d (no NVLink / P2P): use qnb_comm_init (NCCL)", h->device, b.device, p);
        void *q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, b.mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail("cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
        h->peer_base[p] = q;
        c.base[p] = (double *)q;
    }
    h->p2p = c;
    h->rank = rank; h->nranks = nranks;
    drop_graphs(h);
    return 0;
}

int qnb_comm_status(qnb_handle *h) {
    if (!h) return fail("null handle");
    if (!p2p_ready(h)) return 0;
    CU(cudaSetDevice(h->device));
    unsigned ctl[3] = {0, 0, 0};
    CU(cudaMemcpy(ctl, h->arena + h->arena_ctl_off, sizeof ctl, cudaMemcpyDeviceToHost));
    if (ctl[2]) return fail("peer-memory all-reduce: a rank did not arrive at a barrier within 4 s (epoch %u)", ctl[0]);
    return 0;
}

// ---- pipe peaks measured on the spot (the roofline denominators of the force kernels)
int qnb_bench_peak(int device, int which, float *tflops_out) {
    if (!tflops_out) return fail("null argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    void *buf;
    CU(cudaMalloc(&buf, 64));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 6; rep++) {
        CU(cudaEventRecord(e0));
        if (which == 0) k_fma_peak<float><<<blocks, threads>>>((float *)buf, iters, 1.0001f, 0.5f);
        else k_fma_peak<double><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 8.0 * iters * (double)blocks * threads;
        const float tf = (float)(flop / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    CU(cudaGetLastError());
    *tflops_out = best;
    return 0;
}

// md_run's inner loop on device-resident coordinates: lists every nbcycle steps (md.f90:1661), one nonbonded
// evaluation per step; CUDA-event time of the whole loop.
int qnb_bench_md(qnb_handle *h, const double *lambda, int flags, int steps, int nbcycle, float *ms_out) {
    if (!h || !lambda || !ms_out) return fail("null argument");
    if (!h->lists_built) return fail("pair lists have not been built");
    CU(cudaSetDevice(h->device));
    for (int k = 0; k < h->T.s.nstates; k++) h->hlam[k] = lambda[k];
    CU(cudaMemcpyAsync(h->lam_dev, h->hlam, h->T.s.nstates * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));
    CU(cudaEventRecord(h->ev0, h->st));
    for (int k = 0; k < steps; k++) {
        // the step counter runs on across calls (md.f90:1661, mod(istep, NBcycle) == 0): a timed window of 20 steps holds
        // 0.8 list builds on average, not one
        if (nbcycle > 0 && h->md_phase++ % nbcycle == 0) {
            if (build_device(h, h->hx)) return 1;   // as qnb_build_lists: Q partner lists only when the cut-offs ask for it
        }
        if (step_device(h, flags)) return 1;
    }
    CU(cudaEventRecord(h->ev1, h->st));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    CU(cudaGetLastError());
    return 0;
}

static int flush_l2(qnb_handle *h) {
    const size_t bytes = 256u << 20;   // > 126 MB L2
    if (h->flush.ensure(bytes)) return 1;
    CU(cudaMemsetAsync(h->flush.p, 1, bytes, h->st));
    return 0;
}

int qnb_bench_nonbond(qnb_handle *h, const double *lambda, int flags, int steps, int do_flush, float *ms_out) {
    if (!h || !lambda || !ms_out) return fail("null argument");
    if (!h->lists_built) return fail("pair lists have not been built");
    CU(cudaSetDevice(h->device));
    for (int k = 0; k < h->T.s.nstates; k++) h->hlam[k] = lambda[k];
    CU(cudaMemcpyAsync(h->lam_dev, h->hlam, h->T.s.nstates * sizeof(double), cudaMemcpyHostToDevice, h->st));
    float total = 0.f;
    if (do_flush) {
        for (int k = 0; k < steps; k++) {
            if (flush_l2(h)) return 1;
            CU(cudaEventRecord(h->ev0, h->st));
            if (step_device(h, flags)) return 1;
            CU(cudaEventRecord(h->ev1, h->st));
            CU(cudaEventSynchronize(h->ev1));
            float ms;
            CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            total += ms;
        }
    } else {
        CU(cudaStreamSynchronize(h->st));
        CU(cudaEventRecord(h->ev0, h->st));
        for (int k = 0; k < steps; k++)
            if (step_device(h, flags)) return 1;
        CU(cudaEventRecord(h->ev1, h->st));
        CU(cudaEventSynchronize(h->ev1));
        CU(cudaEventElapsedTime(&total, h->ev0, h->ev1));
    }
    CU(cudaGetLastError());
    *ms_out = total;
    return 0;
}

int qnb_bench_build_lists(qnb_handle *h, int reps, float *ms_out) {
    if (!h || !ms_out) return fail("null argument");
    if (!h->lists_built) return fail("pair lists have not been built");
    CU(cudaSetDevice(h->device));
    float total = 0.f;
    h->time_build = true;
    for (int k = 0; k < 3; k++) h->build_ms[k] = 0.f;
    for (int k = 0; k < reps; k++) {
        CU(cudaEventRecord(h->ev0, h->st));
        if (build_device(h, h->hx)) { h->time_build = false; return 1; }
        CU(cudaEventRecord(h->ev1, h->st));
        CU(cudaEventSynchronize(h->ev1));
        float ms;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        total += ms;
        for (int j = 0; j < 3; j++) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, h->ev_bt[2 * j], h->ev_bt[2 * j + 1]) == cudaSuccess) h->build_ms[j] += t / reps;
            else cudaGetLastError();   // this part did not run in the build (no LRF, no units)
        }
    }
    h->time_build = false;
    *ms_out = total;
    return 0;
}

int qnb_bench_last_build_timing(qnb_handle *h, float out[3]) {
    if (!h || !out) return fail("null argument");
    for (int k = 0; k < 3; k++) out[k] = h->build_ms[k];
    return 0;
}

int qnb_bench_kernels(qnb_handle *h, const double *lambda, int flags, int reps, int do_flush, char *names_out,
                      int names_cap, float *ms_out, int ms_cap) {
    if (!h || !lambda || !names_out || !ms_out) { fail("null argument"); return -1; }
    if (!h->lists_built) { fail("pair lists have not been built"); return -1; }
    if (cudaSetDevice(h->device) != cudaSuccess) { fail("cudaSetDevice"); return -1; }
    for (int k = 0; k < h->T.s.nstates; k++) h->hlam[k] = lambda[k];
    cudaMemcpyAsync(h->lam_dev, h->hlam, h->T.s.nstates * sizeof(double), cudaMemcpyHostToDevice, h->st);
    std::string names;
    int n = 0;
    if (h->npk > 0)
        LAUNCH(h, k_pack_step, cdiv(h->npk, 256), 256, 0, h->npk, h->fix, h->D.nat_solute, h->pk_atom.p, h->pk_sw.p, h->pk_q.p, h->pk_ct.p,
               h->x.p, h->px.p, h->py.p, h->pz.p, h->rec_i.p, h->rec_f.p, h->wT.p, h->wown.p, h->wd.p, (double *)nullptr, 0, (const double *)nullptr, (double *)nullptr, 0);
    for (int k = 0; k < K_COUNT; k++) {
        if (!step_kernel_active(h, k, flags)) continue;
        if (n >= ms_cap) break;
        float total = 0.f;
        for (int r = 0; r < reps + 1; r++) {   // first repetition is warm-up
            cudaMemsetAsync(h->out.p, 0, h->nout * sizeof(double), h->st);
            if (do_flush && flush_l2(h)) return -1;
            cudaEventRecord(h->ev0, h->st);
            launch_step_kernel(h, k, h->st, flags);
            cudaEventRecord(h->ev1, h->st);
            cudaEventSynchronize(h->ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev0, h->ev1);
            if (r > 0) total += ms;
        }
        ms_out[n++] = total / reps;
        if (!names.empty()) names += "\n";
        names += kStepKernelNames[k];
    }
    if (cudaGetLastError() != cudaSuccess) { fail("kernel launch failed"); return -1; }
    snprintf(names_out, names_cap, "%s", names.c_str());
    return n;
}

// the all-reduce of [d | E | EQ] alone (gather_nonbond + master sum), on the handle's stream
int qnb_bench_allreduce(qnb_handle *h, int reps, float *ms_out) {
    if (!h || !ms_out) return fail("null argument");
    if (!h->comm && !p2p_ready(h)) return fail("qnb_bench_allreduce: no communicator (qnb_comm_init / qnb_comm_ipc_attach)");
    CU(cudaSetDevice(h->device));
    for (int k = 0; k < reps + 2; k++) {
        if (k == 2) CU(cudaEventRecord(h->ev0, h->st));
        if (allreduce_arena(h, 0, h->nout, h->st)) return 1;
    }
    CU(cudaEventRecord(h->ev1, h->st));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    *ms_out /= (float)std::max(reps, 1);
    return 0;
}

// phases of the last peer-memory all-reduce on this rank, microseconds: waiting for all ranks, summing and delivering the own
// slice, waiting for all ranks to finish
int qnb_bench_allreduce_phases(qnb_handle *h, double out[3]) {
    if (!h || !out) return fail("null argument");
    if (!p2p_ready(h)) return fail("qnb_bench_allreduce_phases: peer-memory all-reduce not attached");
    CU(cudaSetDevice(h->device));
    unsigned long long st[4];
    CU(cudaMemcpy(st, reinterpret_cast<unsigned *>(h->arena + h->arena_ctl_off) + 64, sizeof st, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) out[k] = (double)(st[k + 1] - st[k]) * 1e-3;
    return 0;
}

int64_t qnb_launch_count(qnb_handle *h) { return h ? h->launches : 0; }

int qnb_last_copy_bytes(qnb_handle *h, int64_t *h2d, int64_t *d2h) {
    if (!h) return fail("null handle");
    if (h2d) *h2d = h->last_h2d;
    if (d2h) *d2h = h->last_d2h;
    return 0;
}

int qnb_finalize(qnb_handle *h) {
    if (!h) return 0;
    {
        std::lock_guard<std::mutex> lock(g_plan_mu);   // a cached batch plan must not outlive one of its handles
        for (qnb_handle *p : g_plan.hs) if (p == h) { g_plan.valid = false; g_plan.hs.clear(); break; }
        for (qnb_handle *p : g_parent.hs) if (p == h) { g_parent.valid = false; g_parent.hs.clear(); break; }
    }
    cudaSetDevice(h->device);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->st) cudaStreamSynchronize(h->st);
    drop_graphs(h, true);
    for (int k = 0; k < kAux; k++) { if (h->aux[k]) cudaStreamDestroy(h->aux[k]); if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]); }
    if (h->lrf_st) cudaStreamDestroy(h->lrf_st);
    if (h->ev_lrf) cudaEventDestroy(h->ev_lrf);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_pack) cudaEventDestroy(h->ev_pack);
    for (int k = 0; k < 6; k++) if (h->ev_bt[k]) cudaEventDestroy(h->ev_bt[k]);
    h->crg.release(); h->ljd.release(); h->crgf.release(); h->ljf.release(); h->ctype.release(); h->grp_of_atom.release();
    h->g_first.release(); h->g_n.release(); h->g_switch.release(); h->g_atoms.release(); h->g_nq.release();
    h->u_sw.release(); h->u_grp.release(); h->sp_off.release(); h->sp_partner.release(); h->gs_off.release();
    h->gs_atoms.release(); h->nq_off.release(); h->nq_atoms.release(); h->iqseq.release(); h->is_q.release(); h->excl.release(); h->qbonded.release();
    h->u_excl.release(); h->ljcode.release(); h->sp_code.release(); h->qp_tab.release(); h->qw_tab.release(); h->qp_tabf.release(); h->qw_tabf.release();
    h->qstatic.release(); h->x.release(); h->out.release(); h->lrf.release(); h->upos.release();
    h->cell_of.release(); h->cell_count.release(); h->cell_start.release(); h->cell_items.release(); h->counts.release();
    h->row_tot.release(); h->row_off.release(); h->flag.release(); h->pos.release(); h->flag2.release(); h->pos2.release(); h->qp_list.release();
    h->qw_list.release(); h->qp_shift_atom.release(); h->rows.release(); h->flush.release();
    h->pk_atom.release(); h->pk_ct.release(); h->pk_q.release(); h->pk_qd.release(); h->px.release(); h->py.release(); h->pz.release();
    h->nch.release(); h->choff.release(); h->ucost.release(); h->cost_off.release(); h->wstart_w.release(); h->wstart_s.release(); h->wdesc.release(); h->wrow.release(); h->sdesc.release(); h->srow.release(); h->sspec.release(); h->x_saved.release(); h->lrf_saved.release(); h->wp_theta.release(); h->wp_shell_n.release(); h->wp_shell_list.release(); h->wp_shell_theta.release(); h->lrf_mom.release();
    h->shk_first.release(); h->shk_ij.release(); h->shk_d2.release(); h->shk_winv.release(); h->shk_x.release();
    h->shk_xx.release(); h->shk_iter.release();
    h->bead_atoms.release(); h->bead_base.release(); h->bead_disp.release(); h->bead_eq.release();
    h->item_posf.release(); h->item_scr.release(); h->srcf.release(); h->lrf_planes.release();
    h->upk.release(); h->pk_sw.release(); h->e_cnt.release(); h->e_off.release(); h->rec_i.release(); h->rec_f.release(); h->wT.release(); h->wown.release(); h->wd.release();
    h->pw12.release(); h->ljp.release(); h->pw0.release(); h->ww_pairs.release(); h->pp_pairs.release(); h->pw_pairs.release();
    h->item_pos.release(); h->src.release(); h->cell_unsorted.release(); h->item_nq.release(); h->src_off.release();
    release_host_buffers(h);
    for (int k = 0; k < kMaxPeers; k++) if (h->peer_base[k]) cudaIpcCloseMemHandle(h->peer_base[k]);
    if (h->arena) cudaFree(h->arena);
    if (h->hx) cudaFreeHost(h->hx);
    if (h->hout) cudaFreeHost(h->hout);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return 0;
}

}  // extern "C"
