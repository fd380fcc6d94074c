// TEST INFRASTRUCTURE (CPU suite): C entry points over the PRODUCT's host-side table builder,
// q6_b200/csrc/qnb_tables.hpp (HostTables::build = the library's statement of precompute_interactions and nbqqlist),
// so that tests/test_host_tables_cpu.py can compare it with the oracle without a GPU.  Compiled by the test with g++;
// nothing here is shipped.
#include <cstring>

#include "../../q6_b200/csrc/qnb_tables.hpp"

using namespace qnb;

static void put(const QPar &p, double *out) {
    out[0] = p.A; out[1] = p.B; out[2] = p.el; out[3] = p.score;
}

extern "C" {

void *ht_build(const qnb_system *sys, char *err, int cap) {
    HostTables *T = new HostTables;
    if (!T->build(sys)) {
        if (err && cap > 0) { std::strncpy(err, T->error.c_str(), cap - 1); err[cap - 1] = 0; }
        delete T;
        return nullptr;
    }
    return T;
}
void ht_free(void *h) { delete (HostTables *)h; }

// solute-solute pair (0-based atoms): 0 = excluded, 1 = listed if set; out = vdWA vdWB elec score
int ht_pp(void *h, int i, int j, double *out, int *set);
int ht_qp(void *h, int iq, int state, int atom, double *out);
int ht_pp(void *h, int i, int j, double *out, int *set) {
    QPar p{};
    bool st = false;
    if (!((HostTables *)h)->pp_params(i, j, p, st)) return 0;
    put(p, out);
    *set = st;
    return 1;
}
// batched: ij[2n] 1-based atoms as the lists hold them; ok[n] = listed (not excluded), set[n], out[4n]
void ht_pp_many(void *h, long n, const int *ij, double *out, int *set, int *ok) {
    for (long k = 0; k < n; k++) ok[k] = ht_pp(h, ij[2 * k] - 1, ij[2 * k + 1] - 1, out + 4 * k, set + k);
}
void ht_pw(void *h, int i, int site, double *out) { put(((HostTables *)h)->pw_params(i, site), out); }
void ht_ww(void *h, int a, int b, double *out) {
    HostTables *T = (HostTables *)h;
    put(T->ww_par[(size_t)a * T->s.solv_atom + b], out);
}
int ht_qp(void *h, int iq, int state, int atom, double *out) {   // all 0-based; returns %set
    HostTables *T = (HostTables *)h;
    size_t k = ((size_t)iq * T->s.nstates + state) * T->s.nat_solute + atom;
    put(T->qp_tab[k], out);
    return T->qp_set[k];
}
// batched: ij[2n] = (Q-atom number, topology atom), both 1-based
void ht_qp_many(void *h, long n, const int *ij, int state, double *out, int *set) {
    for (long k = 0; k < n; k++) set[k] = ht_qp(h, ij[2 * k] - 1, state, ij[2 * k + 1] - 1, out + 4 * k);
}
void ht_qw(void *h, int iq, int state, int site, double *out) {
    HostTables *T = (HostTables *)h;
    put(T->qw_tab[((size_t)iq * T->s.nstates + state) * T->s.solv_atom + site], out);
}
long ht_static_count(void *h, int which) {
    HostTables *T = (HostTables *)h;
    return (long)(which == 0 ? T->qq_list.size() : T->qqp_list.size());
}
// ids: topology i, j (0-based), iq, jq (1-based; jq = 0 for qqp), state (0-based), soft
void ht_static(void *h, int which, long k, int *ids, double *out) {
    HostTables *T = (HostTables *)h;
    const StaticQPair &e = which == 0 ? T->qq_list[k] : T->qqp_list[k];
    ids[0] = e.i; ids[1] = e.j; ids[2] = e.iq; ids[3] = e.jq; ids[4] = e.state; ids[5] = e.soft;
    put(e.p, out);
}
int ht_atom_flags(void *h, int atom) {   // bit 0: Q-atom, bit 1: within three bonds of a Q-atom in some state, bit 2: excluded
    HostTables *T = (HostTables *)h;
    return (T->is_q[atom] ? 1 : 0) | (atom < (int)T->qbonded.size() && T->qbonded[atom] ? 2 : 0) | (T->excl[atom] ? 4 : 0);
}

}  // extern "C"
