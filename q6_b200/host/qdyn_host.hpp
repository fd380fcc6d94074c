// qdyn_host.hpp -- the host side above the C ABI (include/qnb.h), in C++ because the reference host is compiled
// Fortran and no Fortran compiler exists in this image.
//
// It mirrors, with the reference's names, argument meaning and error behaviour, what Qdyn6 does on the host before
// and around the nonbonded path (everything here runs on the CPU in the reference too):
//
//   topo_read              topo.f90:522-1114        Q topology file (unchanged format)
//   qatom_load_fep         qatom.f90:322-1050, 1865-1996   FEP/EVB strategy file (unchanged format)
//   prep_sim               simprep.f90:4521-4571 (topology: ljcod, sqrt(eps)), 997-1182 (get_fep),
//                          3592-3721 (charge scaling), nonbondene.f90:3087-3187 (make_qconn)
//   distribute_nonbonds    nonbondene.f90:80-505    i-range assignment (equal shares)
//   init_constraints/shake simprep.f90:2167-2345, bondene.f90:1025-1150   initial solvent SHAKE of qdyn.f90:133
//                          (host code outside the path; kept so that a run starts from the reference's step-0
//                          coordinates)
//   Nonbonded              the module procedures whose bodies become C-ABI calls:
//                          make_pair_lists(Rq,Rcq2,RcLRF2,Rcpp2,Rcpw2,Rcww2)   nonbondene.f90:749
//                          pot_energy_nonbonds(E_loc,EQ_loc,md)                potene.f90:320 (+ nonbond_qq/qqp L176-177)
//   write_out              qalloc.f90:674-806       the nonbonded rows of the energy summary, same formats
//
// There is no CPU fallback: Nonbonded's constructor dies when libqnb finds no usable GPU.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qnb.h"

namespace qdyn {

// die() of qalloc.f90:609-668: the reference prints the message and stops; here the message travels in an exception
// that main() (or the C API below) turns into the same text.
struct Die : std::runtime_error {
    explicit Die(const std::string &m) : std::runtime_error(m) {}
};

constexpr int MAX_NBR_RANGE = 25;  // topo.f90:37
constexpr int NLJTYP = 3;          // avdw/bvdw codes 1..3

struct Topology {
    std::string title;
    double version = 2.0;
    int nat_pro = 0, nat_solute = 0, solv_atom = 3, nwat = 0;
    std::vector<double> xtop;          // [nat_pro][3]
    std::vector<int32_t> iac;          // [nat_pro]
    int nbonds = 0, nbonds_solute = 0;
    std::vector<int32_t> bnd;          // [nbonds][3] i, j, cod
    std::vector<double> bondlib;       // [nbndcod][2] fk, bnd0
    int nangles = 0, nangles_solute = 0, ntors = 0, ntors_solute = 0, nimps = 0, nimps_solute = 0;
    std::vector<double> crg;           // [nat_pro], unscaled
    int ncgp = 0, ncgp_solute = 0, iuse_switch_atom = 1;
    std::vector<int32_t> cgp;          // [ncgp][3] iswitch, first, last
    std::vector<int32_t> cgpatom;
    int natyps = 0, ivdw_rule = 1;
    double el14_scale = 1.0, coulomb_constant = 332.0;
    std::vector<double> iaclib;        // [natyps][7] mass, avdw(3), bvdw(3)
    std::vector<int32_t> lj2;          // [nlj2][2]
    std::vector<int32_t> list14, listex;         // [nat_solute][MAX_NBR_RANGE]
    std::vector<int32_t> list14long, listexlong; // [n][2]
    int nres = 0, nres_solute = 0;
    std::vector<int32_t> res_start;
    std::vector<std::string> res_name;
    int nmol = 0;
    std::vector<int32_t> istart_mol;
    std::vector<std::string> tac;
    int solvent_type = 0;
    bool use_PBC = false;
    double boxlength[3] = {0, 0, 0}, boxcentre[3] = {0, 0, 0};
    double rexcl_o = 0.0, rwat = 0.0, xpcent[3] = {0, 0, 0}, xwcent[3] = {0, 0, 0};
    std::vector<int32_t> excl;         // [nat_pro] 0/1
};

Topology topo_read(const std::string &path);

struct Fep {
    int nstates = 1, nqat = 0, offset = 0;
    bool qq_use_library_charges = false, softcore_use_max_potential = false, qvdw_flag = false;
    int qswitch = 0, nqlib = 0;
    std::vector<int32_t> iqseq;        // [nqat]
    std::vector<double> qcrg;          // [nqat][nstates], unscaled
    std::vector<std::string> qtac;
    std::vector<double> qavdw, qbvdw;  // [nqlib][3]
    std::vector<int32_t> qiac;         // [nqat][nstates]
    std::vector<int32_t> iqexpnb, jqexpnb;
    std::vector<int32_t> el_scale_iq, el_scale_jq;
    std::vector<double> el_scale;      // [n][nstates]
    std::vector<int32_t> exspec_ij;    // [n][2]
    std::vector<int32_t> exspec_flag;  // [n][nstates]
    std::vector<int32_t> qbnd_ij;      // [n][2]
    std::vector<int32_t> qbnd_cod;     // [n][nstates]
    std::vector<double> alpha_max;     // [nqat][nstates]
    std::vector<double> sc_lookup;     // [nqat][natyps+nqat][nstates]
};

// nstates_expected: the [lambdas] count of the md input (checked like qatom.f90:356-378); < 0 = trust the file
Fep qatom_load_fep(const std::string &path, const Topology &topo, int nstates_expected = -1);

// make_qconn / find_bonded (nonbondene.f90:3087-3187); result [iq][atom][state] == Fortran (state,atom,iq)
std::vector<int32_t> make_qconn(int nstates, int nat_solute, int nqat, const std::vector<int32_t> &iqseq,
                                const std::vector<int32_t> &iqatom, const std::vector<int32_t> &bnd_solute,
                                const std::vector<int32_t> &qbnd_ij, const std::vector<int32_t> &qbnd_cod,
                                const std::vector<int32_t> &exspec_ij, const std::vector<int32_t> &exspec_flag);

// The static tables of one node in the Fortran host's layouts (what qnb_system points into).
struct System {
    System() = default;
    // s holds raw pointers into the vectors below: moving keeps them valid (the heap buffers travel), copying would not
    System(const System &) = delete;
    System &operator=(const System &) = delete;
    System(System &&) = default;
    System &operator=(System &&) = default;

    qnb_system s{};                    // scalars filled in; pointers set by view()
    double boxlength[3] = {0, 0, 0};
    std::vector<double> xtop, mass;    // topology coordinates, per-atom masses (for SHAKE)
    std::vector<int32_t> cgp, cgpatom, excl, iqatom, iqseq, iac, ljcod, listex, list14, listexlong, list14long;
    std::vector<double> crg, iaclib;
    std::vector<double> qcrg, qavdw, qbvdw, sc_lookup, el_scale;
    std::vector<int32_t> qiac, iqexpnb, jqexpnb, el_scale_iq, el_scale_jq, qconn;
    // SHAKE constraints of the default input (solvent bonds to hydrogens): rows i, j (1-based) and dist2
    std::vector<int32_t> const_ij;
    std::vector<double> const_dist2;
    std::vector<int32_t> istart_mol;

    // pointers of s refreshed to the vectors above (call after any vector changed)
    const qnb_system *view();
    // calculation_assignment for numnodes == 1 (nonbondene.f90:126-152)
    void full_shard();
    // rank's share of nranks (distribute_nonbonds with equal shares)
    void shard(int rank, int nranks);
};

// topology + get_fep + prep_sim + make_qconn + init_constraints for one node
System prep_sim(const Topology &topo, const Fep *fep, bool use_LRF);

// deep copy of tables that some other host already prepared (sizes from the struct's scalars)
System system_from_struct(const qnb_system &src, const double boxlength[3]);

// contiguous 1-based inclusive i-ranges balanced by per-i counts (nonbondene.f90:171-296, equal shares)
std::vector<std::pair<int, int>> distribute_nonbonds(const std::vector<double> &per_i_counts, int nranks);

// const_mol(:) of init_constraints: first constraint (0-based) of every constrained molecule, plus the end (CSR); the
// topology lists bonds molecule by molecule
std::vector<int32_t> constraint_molecules(const System &sys);
// shake(xx, x) of bondene.f90:1069-1150 over the system's constraints; returns the summed iterations / nmol
int shake(const System &sys, const double *xx, double *x);
// initial_constraint (bondene.f90:1025-1048), coordinate part: xx = x; shake(xx, x)
int initial_constraint(const System &sys, double *x);

// nrgy.f90:37-40, 48-54, 97-110
struct NB_ENERGIES {
    double el = 0, vdw = 0;
};
struct ENERGIES {  // the members the nonbonded path writes
    NB_ENERGIES pp, pw, ww;
    double LRF = 0;
};
struct OQ_ENERGIES {
    double lambda = 0;
    NB_ENERGIES qq, qp, qw;
};

// The nonbonded module procedures of Qdyn6 with their bodies replaced by C-ABI calls.  x, d and EQ(:)%lambda are
// module globals in the reference; here they are public members.
class Nonbonded {
  public:
    Nonbonded(System &sys, int device);
    ~Nonbonded();
    Nonbonded(const Nonbonded &) = delete;
    Nonbonded &operator=(const Nonbonded &) = delete;

    std::vector<double> x, d;   // [3*natom]
    double RcLRF = -1.0;        // the unsquared global the box builders test (nonbondene.f90:3066)
    int64_t nb_pairs[8] = {0};  // nbpp_pair, nbpw_pair, nbww_pair, nbqp_pair, nbqw_pair, nb??_cgp_pair
    // the reference prints the list sizes at a list update only when built with -DDUMP (md.f90:1692-1696); counting
    // the device lists costs host round trips, so nb_pairs is refreshed only while this is set
    bool dump = false;

    // nonbondene.f90:749, the reference's argument list
    void make_pair_lists(double Rq, double Rcq2, double RcLRF2, double Rcpp2, double Rcpw2, double Rcww2);
    // potene.f90:320 (+ the static nbqq/nbqqp terms of potene.f90:176-177): d is ADDED to, E_loc / EQ_loc receive
    // this call's sums; EQ_loc(:)%lambda is read
    void pot_energy_nonbonds(ENERGIES &E_loc, std::vector<OQ_ENERGIES> &EQ_loc, bool md);
    // md.f90 MC_volume / put_back_in_box
    void update_box(const double boxlength[3]);
    // MC_volume keeping / putting back the lists of the current box (md.f90:2022-2058, 2214-2256)
    void save_lists();
    void restore_lists();
    // restrain_solvent + watpol inside the device step (potene.f90:161-167): parameters once, theta_corr when it changes;
    // while restraints_on is set pot_energy_nonbonds(md=.true.) adds both terms to d and last_restraints() returns
    // E%restraint%solvent_radial, E%restraint%water_pol and the per-shell sums the host keeps its averages from
    void set_solvent_restraints(const qnb_solvent_restraints &p);
    void set_theta_corr(const double *theta_corr);
    void last_restraints(double E[2], double *shell_theta_sum, int32_t *shell_n);
    bool restraints_on = false;
    // shake(xx, x) for the solvent (bondene.f90:1069): constraints from System (init_constraints), then per step;
    // xx == nullptr: the coordinates of this step's pot_energy_nonbonds.  Returns the sweeps summed over molecules.
    void set_constraints();
    int64_t shake(const double *xx, double *x);
    // qcp_run's bead loop (qcp.f90:319-372): EQ_out[nbeads][6*nstates]
    void qcp_beads(const double *x_save, int natq, const int32_t *atoms, int nbeads, const double *coord,
                   const double *lambda, double *EQ_out);
    qnb_handle *handle() { return h_; }

  private:
    System &sys_;
    qnb_handle *h_ = nullptr;
};

// The nonbonded rows of write_out (qalloc.f90:674-806) in the reference's formats 6, 26 and 32.
std::string write_out_nonbonded(const System &sys, const ENERGIES &E, const std::vector<OQ_ENERGIES> &EQ, int istep);

}  // namespace qdyn
