// C entry points over qdyn_host (for the ctypes tests and for hosts that are not C++): plain pointers and sizes.
#include <cstring>
#include <memory>
#include <string>

#include "qdyn_host.hpp"

using namespace qdyn;

struct qhost {
    System sys;
    std::unique_ptr<Nonbonded> nb;
    std::string text;
};

static thread_local std::string g_err;

template <class F>
static int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

extern "C" {

const char *qhost_last_error(void) { return g_err.c_str(); }

// topo_read + qatom_load_fep + prep_sim.  fep may be NULL or "" (no Q-atoms); nstates_expected < 0 = trust the file.
int qhost_open(const char *top, const char *fep, int use_lrf, int nstates_expected, qhost **out) {
    return guarded([&] {
        Topology t = topo_read(top);
        std::unique_ptr<qhost> h(new qhost);
        if (fep && fep[0]) {
            Fep f = qatom_load_fep(fep, t, nstates_expected);
            h->sys = prep_sim(t, &f, use_lrf != 0);
        } else {
            h->sys = prep_sim(t, nullptr, use_lrf != 0);
        }
        h->sys.view();
        *out = h.release();
    });
}

// a host object over tables prepared elsewhere (deep copy; boxlength may be NULL)
int qhost_from_system(const qnb_system *src, const double *boxlength, qhost **out) {
    return guarded([&] {
        std::unique_ptr<qhost> h(new qhost);
        h->sys = system_from_struct(*src, boxlength);
        h->sys.view();
        *out = h.release();
    });
}

const qnb_system *qhost_system(qhost *h) { return h->sys.view(); }
const double *qhost_xtop(qhost *h) { return h->sys.xtop.data(); }
const double *qhost_boxlength(qhost *h) { return h->sys.boxlength; }
int qhost_shard(qhost *h, int rank, int nranks) {
    return guarded([&] { h->sys.shard(rank, nranks); });
}
int64_t qhost_constraint_count(qhost *h) { return (int64_t)h->sys.const_dist2.size(); }
// constraint_molecules(): writes up to cap entries of the CSR array, returns its length (molecules + 1)
int64_t qhost_constraint_molecules(qhost *h, int32_t *first, int64_t cap) {
    const std::vector<int32_t> f = constraint_molecules(h->sys);
    for (int64_t k = 0; k < (int64_t)f.size() && k < cap; k++) first[k] = f[k];
    return (int64_t)f.size();
}

// initial_constraint (bondene.f90:1025), coordinate part, in place on x[3*natom]; iterations per molecule in *niter
int qhost_initial_constraint(qhost *h, double *x, int *niter) {
    return guarded([&] {
        int n = initial_constraint(h->sys, x);
        if (niter) *niter = n;
    });
}

// Nonbonded on GPU `device` (dies without one: no CPU path)
int qhost_attach_gpu(qhost *h, int device) {
    return guarded([&] { h->nb.reset(new Nonbonded(h->sys, device)); });
}

int qhost_make_pair_lists(qhost *h, const double *x, double Rq, double Rcq2, double RcLRF2, double Rcpp2, double Rcpw2,
                          double Rcww2, double RcLRF, int64_t counts[8]) {
    return guarded([&] {
        if (!h->nb) throw Die("qhost_make_pair_lists: no GPU attached");
        std::memcpy(h->nb->x.data(), x, sizeof(double) * h->nb->x.size());
        h->nb->RcLRF = RcLRF;
        h->nb->dump = counts != nullptr;
        h->nb->make_pair_lists(Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2);
        if (counts) std::memcpy(counts, h->nb->nb_pairs, sizeof(int64_t) * 8);
    });
}

// d[3*natom] is overwritten with this call's gradient; E[7], EQ[6*nstates] as qnb_nonbond
int qhost_pot_energy_nonbonds(qhost *h, const double *x, const double *lambda, int md, double *d, double *E, double *EQ) {
    return guarded([&] {
        if (!h->nb) throw Die("qhost_pot_energy_nonbonds: no GPU attached");
        Nonbonded &nb = *h->nb;
        const int ns = h->sys.s.nstates;
        std::memcpy(nb.x.data(), x, sizeof(double) * nb.x.size());
        std::fill(nb.d.begin(), nb.d.end(), 0.0);  // potene.f90:109
        ENERGIES El;
        std::vector<OQ_ENERGIES> EQl(ns);
        for (int s = 0; s < ns; s++) EQl[s].lambda = lambda[s];
        nb.pot_energy_nonbonds(El, EQl, md != 0);
        std::memcpy(d, nb.d.data(), sizeof(double) * nb.d.size());
        const double e[7] = {El.pp.el, El.pp.vdw, El.pw.el, El.pw.vdw, El.ww.el, El.ww.vdw, El.LRF};
        std::memcpy(E, e, sizeof e);
        for (int s = 0; s < ns; s++) {
            double *q = EQ + 6 * s;
            q[0] = EQl[s].qq.el, q[1] = EQl[s].qq.vdw, q[2] = EQl[s].qp.el;
            q[3] = EQl[s].qp.vdw, q[4] = EQl[s].qw.el, q[5] = EQl[s].qw.vdw;
        }
    });
}

// the nonbonded rows of write_out for given energies (E[7], EQ[6*nstates], lambda[nstates]); returns a string owned by h
const char *qhost_write_out(qhost *h, const double *E, const double *EQ, const double *lambda, int istep) {
    ENERGIES El;
    El.pp = {E[0], E[1]};
    El.pw = {E[2], E[3]};
    El.ww = {E[4], E[5]};
    El.LRF = E[6];
    const int ns = h->sys.s.nstates;
    std::vector<OQ_ENERGIES> EQl(ns);
    for (int s = 0; s < ns; s++) {
        const double *q = EQ + 6 * s;
        EQl[s].lambda = lambda[s];
        EQl[s].qq = {q[0], q[1]};
        EQl[s].qp = {q[2], q[3]};
        EQl[s].qw = {q[4], q[5]};
    }
    h->text = write_out_nonbonded(h->sys, El, EQl, istep);
    return h->text.c_str();
}

void qhost_close(qhost *h) { delete h; }

}  // extern "C"
