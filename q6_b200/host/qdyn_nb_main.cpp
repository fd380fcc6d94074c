// qdyn_nb -- the nonbonded part of a Qdyn6 run from the unchanged input files, host side in C++ over the C ABI.
//
//   qdyn_nb <topology> <fep|-> [--q_atom R] [--lrf R] [--solute_solute R] [--solute_solvent R] [--solvent_solvent R]
//           [--lambda l1,l2,..] [--no-lrf] [--no-shake] [--steps N] [--non_bond K] [--device D]
//
// What qdyn.f90:113-160 + md_run do around the path: topo_read, qatom_load_fep, prep_sim, initial_constraint (iseed > 0),
// then make_pair_lists every `non_bond` steps and pot_energy every step (md.f90:1661-1742).  The integrator, bonded
// terms and restraints are not part of this repository, so coordinates stay where the initial SHAKE puts them; the
// program prints the step-0 energy summary in write_out's formats (so `grep Q-surr.` gives the line
// tests/basic_tests/eval_test.sh:97-101 checks) and, with --steps, the host-measured time per step of the path.
// Cut-off defaults are those of md.f90 (10 / 10 / 10, q_atom 99, lrf 99).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "qdyn_host.hpp"

using namespace qdyn;

static std::vector<double> parse_list(const char *s) {
    std::vector<double> out;
    const char *p = s;
    while (*p) {
        char *end = nullptr;
        double v = std::strtod(p, &end);
        if (end == p) break;
        out.push_back(v);
        p = end;
        while (*p == ',' || *p == ' ') p++;
    }
    return out;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <topology> <fep|-> [--q_atom R] [--lrf R] [--solute_solute R] [--solute_solvent R] "
                             "[--solvent_solvent R] [--lambda l1,l2] [--no-lrf] [--no-shake] [--steps N] [--non_bond K] "
                             "[--device D]\n", argv[0]);
        return 2;
    }
    double Rcq = 99.0, RcLRF = 99.0, Rcpp = 10.0, Rcpw = 10.0, Rcww = 10.0;
    bool use_lrf = true, do_shake = true;
    int steps = 0, nbcycle = 25, device = 0;
    std::vector<double> lambda;
    for (int a = 3; a < argc; a++) {
        std::string k = argv[a];
        auto val = [&]() -> const char * {
            if (a + 1 >= argc) {
                std::fprintf(stderr, ">>> ERROR: %s needs a value\n", k.c_str());
                std::exit(2);
            }
            return argv[++a];
        };
        if (k == "--q_atom") Rcq = std::atof(val());
        else if (k == "--lrf") RcLRF = std::atof(val());
        else if (k == "--solute_solute") Rcpp = std::atof(val());
        else if (k == "--solute_solvent") Rcpw = std::atof(val());
        else if (k == "--solvent_solvent") Rcww = std::atof(val());
        else if (k == "--lambda") lambda = parse_list(val());
        else if (k == "--no-lrf") use_lrf = false;
        else if (k == "--no-shake") do_shake = false;
        else if (k == "--steps") steps = std::atoi(val());
        else if (k == "--non_bond") nbcycle = std::atoi(val());
        else if (k == "--device") device = std::atoi(val());
        else {
            std::fprintf(stderr, ">>> ERROR: unknown option %s\n", k.c_str());
            return 2;
        }
    }
    try {
        Topology topo = topo_read(argv[1]);
        System sys;
        if (std::strcmp(argv[2], "-") != 0) {
            Fep fep = qatom_load_fep(argv[2], topo, lambda.empty() ? -1 : (int)lambda.size());
            sys = prep_sim(topo, &fep, use_lrf);
        } else {
            sys = prep_sim(topo, nullptr, use_lrf);
        }
        sys.view();
        const int ns = sys.s.nstates;
        if (lambda.empty()) {  // md.f90: lambdas default to state 1
            lambda.assign(ns, 0.0);
            lambda[0] = 1.0;
        }
        std::printf("No. of atoms = %d (solute %d), waters = %d, charge groups = %d, Q-atoms = %d, states = %d, %s\n",
                    sys.s.natom, sys.s.nat_solute, sys.s.nwat, sys.s.ncgp, sys.s.nqat, ns,
                    sys.s.use_PBC ? "periodic box" : "sphere");
        Nonbonded nb(sys, device);
        nb.x = sys.xtop;
        if (do_shake) {
            int it = initial_constraint(sys, nb.x.data());
            std::printf("Initial x-constraint required%4d iterations per molecule on average.\n", it);
        }
        nb.RcLRF = use_lrf ? RcLRF : -1.0;
        const double RcLRF2 = use_lrf ? RcLRF * RcLRF : -1.0;
        nb.dump = true;  // list sizes for the log of the first update only
        nb.make_pair_lists(Rcq, Rcq * Rcq, RcLRF2, Rcpp * Rcpp, Rcpw * Rcpw, Rcww * Rcww);
        nb.dump = false;
        std::printf("%-13s%13s%13s%13s%13s\n", "solute-solute", "solute-water", "water-water", "Q-solute", "Q-water");
        std::printf("%13lld%13lld%13lld%13lld%13lld\n", (long long)nb.nb_pairs[0], (long long)nb.nb_pairs[1],
                    (long long)nb.nb_pairs[2], (long long)nb.nb_pairs[3], (long long)nb.nb_pairs[4]);
        ENERGIES E;
        std::vector<OQ_ENERGIES> EQ(ns);
        for (int s = 0; s < ns; s++) EQ[s].lambda = lambda[s];
        std::fill(nb.d.begin(), nb.d.end(), 0.0);
        nb.pot_energy_nonbonds(E, EQ, true);
        std::fputs(write_out_nonbonded(sys, E, EQ, 0).c_str(), stdout);
        if (steps > 0) {
            for (int istep = 0; istep < nbcycle + 3; istep++) {  // warm-up: first rebuild, graph instantiation
                if (istep % nbcycle == 0)
                    nb.make_pair_lists(Rcq, Rcq * Rcq, RcLRF2, Rcpp * Rcpp, Rcpw * Rcpw, Rcww * Rcww);
                std::fill(nb.d.begin(), nb.d.end(), 0.0);
                nb.pot_energy_nonbonds(E, EQ, true);
            }
            auto t0 = std::chrono::steady_clock::now();
            for (int istep = 0; istep < steps; istep++) {
                if (istep % nbcycle == 0)
                    nb.make_pair_lists(Rcq, Rcq * Rcq, RcLRF2, Rcpp * Rcpp, Rcpw * Rcpw, Rcww * Rcww);
                std::fill(nb.d.begin(), nb.d.end(), 0.0);  // potene.f90:109
                nb.pot_energy_nonbonds(E, EQ, true);
            }
            double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            std::printf("nonbonded path: %d steps, list update every %d: %.4f ms per step (host clock, host buffers)\n",
                        steps, nbcycle, sec / steps * 1e3);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\nQ terminated abnormally\n", e.what());  // die(), qalloc.f90:609-668
        return 255;
    }
    return 0;
}
