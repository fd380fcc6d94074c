// qdyn_host.cpp -- see qdyn_host.hpp.  Host-only C++17 (no CUDA); links against libqnb.so for the Nonbonded class.
#include "qdyn_host.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <sstream>

namespace qdyn {

// ------------------------------------------------------------------------------------------------ text helpers
namespace {

struct BadValue : std::runtime_error {
    explicit BadValue(const std::string &m) : std::runtime_error(m) {}
};

std::string strip(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

std::string rstrip(const std::string &s) {
    size_t b = s.size();
    while (b > 0 && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(0, b);
}

std::string lower(std::string s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

std::string upper(std::string s) {
    for (auto &c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}

std::vector<std::string> split_ws(const std::string &s) {
    std::vector<std::string> out;
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && std::isspace((unsigned char)s[i])) i++;
        size_t j = i;
        while (j < n && !std::isspace((unsigned char)s[j])) j++;
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

// first token and the (stripped) rest of the line
std::pair<std::string, std::string> split_first(const std::string &s) {
    std::string t = strip(s);
    size_t j = 0;
    while (j < t.size() && !std::isspace((unsigned char)t[j])) j++;
    return {t.substr(0, j), strip(t.substr(j))};
}

std::string substr_safe(const std::string &s, size_t pos, size_t len) {
    if (pos >= s.size()) return std::string();
    return s.substr(pos, len);
}

bool parse_int(const std::string &tok, long &v) {
    if (tok.empty()) return false;
    char *end = nullptr;
    v = std::strtol(tok.c_str(), &end, 10);
    return end && *end == '\0' && end != tok.c_str();
}

bool parse_float(const std::string &tok0, double &v) {
    if (tok0.empty()) return false;
    std::string tok = tok0;
    for (auto &c : tok) {  // Fortran D exponents
        if (c == 'D' || c == 'd') c = 'E';
    }
    char *end = nullptr;
    v = std::strtod(tok.c_str(), &end);
    return end && *end == '\0' && end != tok.c_str();
}

int to_int(const std::string &tok) {
    long v;
    if (!parse_int(tok, v)) throw BadValue("invalid integer '" + tok + "'");
    return (int)v;
}

double to_float(const std::string &tok) {
    double v;
    if (!parse_float(tok, v)) throw BadValue("invalid number '" + tok + "'");
    return v;
}

std::string commas_to_spaces(std::string s) {
    for (auto &c : s)
        if (c == ',') c = ' ';
    return s;
}

// Fortran sequential formatted input: every READ starts on a fresh record.
class Records {
  public:
    explicit Records(const std::string &text) {
        size_t i = 0;
        while (i <= text.size()) {
            size_t j = text.find('\n', i);
            if (j == std::string::npos) j = text.size();
            std::string ln = text.substr(i, j - i);
            if (!ln.empty() && ln.back() == '\r') ln.pop_back();
            lines_.push_back(ln);
            i = j + 1;
        }
    }
    std::string line() {
        if (pos_ >= lines_.size()) throw BadValue("unexpected end of topology file");
        return lines_[pos_++];
    }
    void skip() { line(); }

    // list-directed read of up to nmax leading values; stops at the first bad token
    static std::vector<long> leading_int(const std::string &line, size_t nmax) {
        std::vector<long> out;
        for (auto &tok : split_ws(commas_to_spaces(line))) {
            if (out.size() == nmax) break;
            long v;
            if (!parse_int(tok, v)) break;
            out.push_back(v);
        }
        return out;
    }
    static std::vector<double> leading_float(const std::string &line, size_t nmax) {
        std::vector<double> out;
        for (auto &tok : split_ws(commas_to_spaces(line))) {
            if (out.size() == nmax) break;
            double v;
            if (!parse_float(tok, v)) break;
            out.push_back(v);
        }
        return out;
    }
    // one list-directed READ of n items (spans records, drops the rest of the last one)
    std::vector<int32_t> ints(size_t n) {
        std::vector<int32_t> out;
        out.reserve(n);
        while (out.size() < n) {
            for (auto &tok : split_ws(commas_to_spaces(line()))) {
                if (out.size() == n) break;
                out.push_back(to_int(tok));
            }
        }
        return out;
    }
    std::vector<double> floats(size_t n) {
        std::vector<double> out;
        out.reserve(n);
        while (out.size() < n) {
            for (auto &tok : split_ws(commas_to_spaces(line()))) {
                if (out.size() == n) break;
                out.push_back(to_float(tok));
            }
        }
        return out;
    }
    // fixed-format READ of n one-character fields, `width` per record (80i1 / 80l1)
    std::string fixed_chars(size_t n, size_t width = 80) {
        std::string out;
        out.reserve(n);
        size_t left = n;
        while (left > 0) {
            std::string rec = line();
            size_t take = std::min(left, width);
            std::string part = rec.substr(0, std::min(take, rec.size()));
            part.resize(take, ' ');
            out += part;
            left -= take;
        }
        return out;
    }

  private:
    std::vector<std::string> lines_;
    size_t pos_ = 0;
};

std::string read_file(const std::string &path, const char *what) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Die(std::string(">>>>> ERROR: Could not open ") + what + " file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

template <class T>
T at(const std::vector<T> &v, size_t i, const char *what) {
    if (i >= v.size()) throw BadValue(std::string("missing value: ") + what);
    return v[i];
}

void read_topology(Records &rec, Topology &t) {
    // 1. header (topo.f90:569-605)
    std::string title = rec.line().substr(0, 80);
    if (strip(title) == "Q topology file") {
        for (;;) {
            auto kv = split_first(rec.line());
            std::string key = upper(kv.first);
            if (key == "TITLE") {
                title = kv.second;
            } else if (key == "VERSION") {
                // the leading "a.b" part of the version string (L583-586)
                auto first = split_ws(kv.second);
                if (first.empty()) throw BadValue("version");
                std::string v = first[0];
                size_t p1 = v.find('.');
                if (p1 != std::string::npos) {
                    size_t p2 = v.find('.', p1 + 1);
                    if (p2 != std::string::npos) v = v.substr(0, p2);
                }
                t.version = to_float(v);
            } else if (key == "END") {
                break;
            }
        }
    } else {
        t.version = 2.0;
    }
    t.title = title;

    // 2. nat_pro, nat_solute, solv_atom (L612-651)
    {
        auto v = Records::leading_int(rec.line(), 4);
        if (v.empty()) throw BadValue("nat_pro");
        t.nat_pro = (int)v[0];
        if (v.size() >= 3) {
            t.nat_solute = (int)v[1];
            t.solv_atom = (int)v[2];
        } else if (v.size() == 2) {
            t.nat_solute = (int)v[1];
            t.solv_atom = 3;
        } else {
            t.nat_solute = t.nat_pro;
            t.solv_atom = 3;
        }
        if (t.solv_atom == 0) {
            t.nwat = 0;
            t.solv_atom = 1;
        } else {
            t.nwat = (t.nat_pro - t.nat_solute) / t.solv_atom;
        }
    }
    // 3. coordinates (L661)
    if (t.nat_pro > 0) t.xtop = rec.floats(3 * (size_t)t.nat_pro);
    // 4. atom codes (L668-669)
    rec.skip();
    if (t.nat_pro > 0) t.iac = rec.ints(t.nat_pro);
    // 5. bonds (L676-707)
    {
        auto v = Records::leading_int(rec.line(), 2);
        t.nbonds = (int)at(v, 0, "nbonds");
        t.nbonds_solute = v.size() > 1 ? (int)v[1] : t.nbonds;
        if (t.nbonds > 0) t.bnd = rec.ints(3 * (size_t)t.nbonds);
        int nbndcod = (int)at(Records::leading_int(rec.line(), 1), 0, "nbndcod");
        t.bondlib.assign(2 * (size_t)nbndcod, 0.0);
        for (int i = 0; i < nbndcod; i++) {
            auto b = Records::leading_float(rec.line(), 3);  // code number, fk, bnd0 [, SYBYL type]
            if (b.size() >= 3) {
                t.bondlib[2 * i] = b[1];
                t.bondlib[2 * i + 1] = b[2];
            }
        }
    }
    // 6. angles (L712-743)
    {
        auto v = Records::leading_int(rec.line(), 2);
        t.nangles = (int)at(v, 0, "nangles");
        t.nangles_solute = v.size() > 1 ? (int)v[1] : t.nangles;
        if (t.nangles > 0) rec.ints(4 * (size_t)t.nangles);
        int n = (int)at(Records::leading_int(rec.line(), 1), 0, "nangcod");
        for (int i = 0; i < n; i++) rec.skip();
    }
    // 7. torsions (L748-769)
    {
        auto v = Records::leading_int(rec.line(), 2);
        t.ntors = (int)at(v, 0, "ntors");
        t.ntors_solute = v.size() > 1 ? (int)v[1] : t.ntors;
        if (t.ntors > 0) rec.ints(5 * (size_t)t.ntors);
        int n = (int)at(Records::leading_int(rec.line(), 1), 0, "ntorcod");
        for (int i = 0; i < n; i++) rec.floats(5);
    }
    // 8. impropers (L773-796)
    {
        auto v = Records::leading_int(rec.line(), 2);
        t.nimps = (int)at(v, 0, "nimps");
        t.nimps_solute = v.size() > 1 ? (int)v[1] : t.nimps;
        if (t.nimps > 0) rec.ints(5 * (size_t)t.nimps);
        int n = (int)at(Records::leading_int(rec.line(), 1), 0, "nimpcod");
        for (int i = 0; i < n; i++) rec.floats(3);
    }
    // 9. charges (L801-804)
    {
        auto v = Records::leading_int(rec.line(), 1);
        if (!v.empty()) t.nat_pro = (int)v[0];
        if (t.nat_pro > 0) t.crg = rec.floats(t.nat_pro);
    }
    // 10. charge groups (L811-832)
    {
        auto v = Records::leading_int(rec.line(), 3);
        t.ncgp = (int)at(v, 0, "ncgp");
        t.ncgp_solute = v.size() > 1 ? (int)v[1] : t.ncgp;
        t.iuse_switch_atom = v.size() > 2 ? (int)v[2] : 1;
        t.cgp.assign(3 * (size_t)t.ncgp, 0);
        int k = 1;
        for (int i = 0; i < t.ncgp; i++) {
            auto h = rec.ints(2);
            int n = h[0], isw = h[1];
            t.cgp[3 * i] = isw;
            t.cgp[3 * i + 1] = k;
            t.cgp[3 * i + 2] = k + n - 1;
            auto a = rec.ints(n);
            t.cgpatom.insert(t.cgpatom.end(), a.begin(), a.end());
            k += n;
        }
    }
    // 11. masses and vdW parameters (L847-891)
    {
        t.natyps = rec.ints(1)[0];
        t.ivdw_rule = rec.ints(1)[0];
        auto v = Records::leading_float(rec.line(), 2);
        t.el14_scale = at(v, 0, "el14_scale");
        t.coulomb_constant = (v.size() > 1 && v[1] > 0) ? v[1] : 332.0;
        t.iaclib.assign(7 * (size_t)t.natyps, 0.0);
        rec.skip();
        auto m = rec.floats(t.natyps);
        for (int i = 0; i < t.natyps; i++) t.iaclib[7 * i] = m[i];
        for (int j = 0; j < NLJTYP; j++) {
            rec.skip();
            auto a = rec.floats(t.natyps);
            for (int i = 0; i < t.natyps; i++) t.iaclib[7 * i + 1 + j] = a[i];
            rec.skip();
            auto b = rec.floats(t.natyps);
            for (int i = 0; i < t.natyps; i++) t.iaclib[7 * i + 4 + j] = b[i];
        }
        int nlj2 = rec.ints(1)[0];
        for (int i = 0; i < nlj2; i++) {
            auto p = rec.ints(2);
            t.lj2.push_back(p[0]);
            t.lj2.push_back(p[1]);
        }
    }
    // 12. 1-4 neighbour and exclusion lists (L912-963)
    {
        rec.ints(1);  // n14nbrs
        if (t.nat_solute > 0) {
            std::string s = rec.fixed_chars((size_t)MAX_NBR_RANGE * t.nat_solute);
            t.list14.resize(s.size());
            for (size_t i = 0; i < s.size(); i++) t.list14[i] = s[i] == '1';
        }
        int n14long = rec.ints(1)[0];
        for (int i = 0; i < n14long; i++) {
            auto p = rec.ints(2);
            t.list14long.push_back(p[0]);
            t.list14long.push_back(p[1]);
        }
        rec.ints(1);  // nexnbrs
        if (t.nat_solute > 0) {
            std::string s = rec.fixed_chars((size_t)MAX_NBR_RANGE * t.nat_solute);
            t.listex.resize(s.size());
            for (size_t i = 0; i < s.size(); i++) t.listex[i] = s[i] == '1';
        }
        int nexlong = rec.ints(1)[0];
        for (int i = 0; i < nexlong; i++) {
            auto p = rec.ints(2);
            t.listexlong.push_back(p[0]);
            t.listexlong.push_back(p[1]);
        }
    }
    // residue / molecule bookkeeping (L966-995)
    {
        std::string line = rec.line();
        if (strip(line).empty()) line = rec.line();
        auto v = Records::leading_int(line, 2);
        t.nres = (int)at(v, 0, "nres");
        t.nres_solute = v.size() > 1 ? (int)v[1] : t.nres;
        if (t.nres > 0) t.res_start = rec.ints(t.nres);
        rec.skip();
        int nrec = t.nres > 0 ? (t.nres + 15) / 16 : 0;
        for (int i = 0; i < nrec; i++) {
            std::string ln = rec.line();
            for (int j = 0; j < std::min(16, t.nres - i * 16); j++) t.res_name.push_back(substr_safe(ln, 5 * j, 4));
        }
        t.nmol = rec.ints(1)[0];
        if (t.nmol > 0) t.istart_mol = rec.ints(t.nmol);
    }
    // atom type names (L997-1031)
    {
        rec.skip();
        int nrec = t.natyps > 0 ? (t.natyps + 7) / 8 : 0;
        for (int i = 0; i < nrec; i++) {
            std::string ln = rec.line();
            for (int j = 0; j < std::min(8, t.natyps - i * 8); j++) t.tac.push_back(strip(substr_safe(ln, 9 * j, 8)));
        }
        rec.skip();
        int nskip = t.natyps > 0 ? (t.natyps + 12) / 13 : 0;
        for (int i = 0; i < nskip; i++) rec.skip();
    }
    t.excl.assign(t.nat_pro, 0);
    if (t.version < 4) return;
    // solvent type and boundary (L1041-1100)
    t.solvent_type = rec.ints(1)[0];
    std::string line = rec.line();
    auto toks = split_ws(line);
    if (!toks.empty() && toks[0] == "PBC") {
        t.use_PBC = true;
        auto b = rec.floats(3), c = rec.floats(3);
        for (int k = 0; k < 3; k++) {
            t.boxlength[k] = b[k];
            t.boxcentre[k] = c[k];
        }
    } else {
        t.use_PBC = false;
        auto v = Records::leading_float(line, 3);
        t.rexcl_o = at(v, 0, "rexcl_o");
        t.rwat = t.version < 5.01 ? at(v, 2, "rwat") : at(v, 1, "rwat");
        auto p = rec.floats(3), w = rec.floats(3);
        for (int k = 0; k < 3; k++) {
            t.xpcent[k] = p[k];
            t.xwcent[k] = w[k];
        }
        rec.ints(2);  // nexats, nexwat
        std::string s = rec.fixed_chars(t.nat_pro);
        for (int i = 0; i < t.nat_pro; i++) t.excl[i] = s[i] == 'T';
    }
}

}  // namespace

Topology topo_read(const std::string &path) {
    Records rec(read_file(path, "topology"));
    Topology t;
    try {
        read_topology(rec, t);
    } catch (const BadValue &e) {  // the reference's err=1000 path (topo.f90:1106)
        throw Die(std::string(">>>>> ERROR: Could not read topology file. (") + e.what() + ")");
    }
    return t;
}

// ------------------------------------------------------------------------------------------------ FEP file
namespace {

// prmfile.f90:845-879
std::string strip_comment(const std::string &line) {
    for (size_t i = 0; i < line.size(); i++) {
        char c = line[i];
        if (c == '!') return rstrip(line.substr(0, i));
        if ((c == '#' || c == '*') && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t'))
            return rstrip(line.substr(0, i));
    }
    return rstrip(line);
}

using Sections = std::map<std::string, std::vector<std::string>>;

Sections parse_sections(const std::string &text) {
    Sections sec;
    std::string cur;
    bool have = false;
    std::istringstream in(text);
    std::string raw;
    while (std::getline(in, raw)) {
        std::string s = strip(raw);
        if (s.empty() || s[0] == '!' || s[0] == '#' || s[0] == '*') continue;
        if (s[0] == '[') {
            size_t e = s.find(']');
            if (e == std::string::npos) throw Die(">>>>> ERROR: malformed section heading in fep file: " + s);
            cur = lower(strip(s.substr(1, e - 1)));
            sec[cur];
            have = true;
            continue;
        }
        s = strip_comment(s);
        if (!s.empty() && have) sec[cur].push_back(s);
    }
    return sec;
}

bool logical(const std::string &s) {
    std::string v = lower(strip(s));
    return v == "on" || v == "true" || v == ".true." || v == "yes" || v == "1" || v == "t";
}

void softcore(const Sections &sec, Fep &f, const Topology &topo) {
    // [softcore] and the sc_lookup table (qatom.f90:1865-1991)
    const int nq = f.nqat, ns = f.nstates, nt = topo.natyps;
    f.sc_lookup.assign((size_t)nq * (nt + nq) * ns, 0.0);
    f.alpha_max.assign((size_t)nq * ns, 0.0);
    auto it = sec.find("softcore");
    if (it == sec.end()) return;
    if (!f.qvdw_flag) throw Die(">>>>> ERROR: Q-atom types must be redefined in \"change_atoms\" section");
    const auto &lines = it->second;
    if ((int)lines.size() != nq) throw Die(">>>>> ERROR: Alpha must be given for every Q-atom!");
    for (auto &ln : lines) {
        auto tok = split_ws(ln);
        int iq = to_int(at(tok, 0, "softcore q-atom"));
        if (iq < 1 || iq > nq) throw Die(">>>>> ERROR: invalid q-atom number in [softcore]");
        for (int s = 0; s < ns; s++) f.alpha_max[(size_t)(iq - 1) * ns + s] = to_float(at(tok, 1 + s, "softcore alpha"));
    }
    const bool geom = topo.ivdw_rule == 1;
    auto am = [&](int i, int s) { return f.alpha_max[(size_t)i * ns + s]; };
    auto sc = [&](int i, int j, int s) -> double & { return f.sc_lookup[((size_t)i * (nt + nq) + j) * ns + s]; };
    for (int i = 0; i < nq; i++) {
        for (int s = 0; s < ns; s++) {
            const int ti = f.qiac[(size_t)i * ns + s] - 1;
            const double aq = f.qavdw[(size_t)ti * NLJTYP], bq = f.qbvdw[(size_t)ti * NLJTYP];
            for (int j = 0; j < nt; j++) {  // q - surroundings
                if (f.softcore_use_max_potential) {
                    // qdyn.f90:114-116: topology() has already square-rooted the library epsilons (simprep.f90:4567-4571)
                    // when get_fep -> qatom_load_fep runs; the Q-atom epsilons (bq) have not been
                    const double aj = topo.iaclib[7 * j + 1];
                    const double bj = geom ? topo.iaclib[7 * j + 4] : std::sqrt(std::fabs(topo.iaclib[7 * j + 4]));
                    if (am(i, s) > 1e-6) {
                        if (geom)
                            sc(i, j, s) = (-bq * bj + std::sqrt(bq * bq * bj * bj + 4.0 * am(i, s) * aq * aj)) /
                                          (2.0 * am(i, s));
                        else
                            sc(i, j, s) = (-2.0 * std::sqrt(bq) * bj +
                                           2.0 * std::sqrt(bq * (bj * bj) + am(i, s) * std::sqrt(bq) * bj)) *
                                          std::pow(aq + aj, 6) / (2.0 * am(i, s));
                    }
                } else {
                    sc(i, j, s) = am(i, s);
                }
            }
            for (int j = 0; j < nq; j++) {  // q - q
                if (am(i, s) > 1e-6 || am(j, s) > 1e-6) {
                    const int tj = f.qiac[(size_t)j * ns + s] - 1;
                    const double aj = f.qavdw[(size_t)tj * NLJTYP], bj = f.qbvdw[(size_t)tj * NLJTYP];
                    if (f.softcore_use_max_potential) {
                        double a = std::min(am(i, s), am(j, s));
                        if (am(i, s) < 1e-6 || am(j, s) < 1e-6) a = std::max(am(i, s), am(j, s));
                        double v;
                        if (geom)
                            v = (-bq * bj + std::sqrt(bq * bq * bj * bj + 4.0 * a * aq * aj)) / (2.0 * a);
                        else
                            v = (-2.0 * std::sqrt(bq * bj) + 2.0 * std::sqrt(bq * bj + a * std::sqrt(bq * bj))) *
                                std::pow(aq + aj, 6) / (2.0 * a);
                        sc(i, nt + j, s) = v;
                    } else {
                        sc(i, nt + j, s) = std::max(am(i, s), am(j, s));
                    }
                }
            }
        }
    }
}

const std::vector<std::string> &section(const Sections &sec, const char *name) {
    static const std::vector<std::string> empty;
    auto it = sec.find(name);
    return it == sec.end() ? empty : it->second;
}

void load_fep_impl(const Sections &sec, const Topology &topo, int nstates_expected, Fep &f) {
    // ---- qatom_load_atoms (qatom.f90:322-640)
    if (!sec.count("atoms")) throw Die(">>> WARNING: No [atoms] section in fep file. Aborting file loading.");
    std::map<std::string, std::string> fepsec;
    for (auto &ln : section(sec, "fep")) {
        auto kv = split_first(ln);
        fepsec[lower(kv.first)] = kv.second;
    }
    f.nstates = fepsec.count("states") ? to_int(fepsec["states"]) : 1;
    if (nstates_expected >= 0 && f.nstates != nstates_expected)
        throw Die(">>>>> ERROR: Mismatch between nstates in input and FEP file");
    if (f.nstates == 0) throw Die(">>>>> ERROR: Number of states must be at least 1. Aborting.");
    f.qq_use_library_charges = fepsec.count("qq_use_library_charges") && logical(fepsec["qq_use_library_charges"]);
    f.softcore_use_max_potential =
        fepsec.count("softcore_use_max_potential") && logical(fepsec["softcore_use_max_potential"]);
    int offset = 0;
    if (fepsec.count("offset")) {
        offset = to_int(fepsec["offset"]);
        if (offset < 1 || offset > topo.nat_solute)
            throw Die(">>>>> ERROR: Invalid topology atom number offset value:" + std::to_string(offset));
    } else if (fepsec.count("offset_name")) {
        std::string name = strip(fepsec["offset_name"].substr(0, 4));
        offset = -1;
        for (int i = 0; i < topo.nres_solute && i < (int)topo.res_name.size(); i++) {
            if (strip(topo.res_name[i]) == name) {
                offset = topo.res_start[i] - 1;
                break;
            }
        }
        if (offset == -1) throw Die(">>>>> ERROR: Residue name " + name + " not found.");
    } else if (fepsec.count("offset_residue")) {
        int r = to_int(fepsec["offset_residue"]);
        if (r < 1 || r > topo.nres_solute)
            throw Die(">>>>> ERROR: Invalid residue number for offset:" + std::to_string(r));
        offset = topo.res_start[r - 1] - 1;
    }
    f.offset = offset;

    const auto &atoms = section(sec, "atoms");
    if (atoms.empty()) return;  // zero Q-atoms: fep file is not loaded (qatom.f90:434-439)
    std::vector<std::pair<int, int>> pairs;
    for (auto &ln : atoms) {
        auto tok = split_ws(ln);
        long a, b;
        if (tok.size() < 2 || !parse_int(tok[0], a) || !parse_int(tok[1], b))
            throw Die("only the 'int int' form of [atoms] is supported by this reader");
        pairs.push_back({(int)a, (int)b});
    }
    f.nqat = 0;
    for (auto &p : pairs) f.nqat = std::max(f.nqat, p.first);
    f.iqseq.assign(f.nqat, 0);
    for (auto &p : pairs) {
        int topno = p.second + offset;
        if (topno < 1 || topno > topo.nat_solute)
            throw Die(">>>>> ERROR: invalid topology atom number " + std::to_string(topno) + " for Q-atom " +
                      std::to_string(p.first));
        if (p.first < 1) throw Die(">>>>> ERROR: invalid Q-atom number " + std::to_string(p.first));
        f.iqseq[p.first - 1] = topno;
    }
    const int nq = f.nqat, ns = f.nstates;

    // ---- get_fep copies the topology charges first (simprep.f90:1027-1031)
    f.qcrg.assign((size_t)nq * ns, 0.0);
    for (int i = 0; i < nq; i++) {
        if (f.iqseq[i] < 1) throw Die(">>>>> ERROR: Q-atom " + std::to_string(i + 1) + " is not assigned in [atoms]");
        for (int s = 0; s < ns; s++) f.qcrg[(size_t)i * ns + s] = topo.crg[f.iqseq[i] - 1];
    }

    // ---- qatom_load_fep
    if (topo.use_PBC) {
        if (!sec.count("pbc")) throw Die(">>>>> ERROR: Section PBC is required when using periodic boundary.");
        std::map<std::string, std::string> kv;
        for (auto &ln : section(sec, "pbc")) {
            auto p = split_first(ln);
            kv[p.first] = p.second;
        }
        if (!kv.count("switching_atom")) throw Die(">>>>> ERROR: Switching atom could not be read.");
        f.qswitch = to_int(kv["switching_atom"]) + offset;
    }

    auto valid_q = [&](int iq) { return iq >= 1 && iq <= nq && f.iqseq[iq - 1] != 0; };

    for (auto &ln : section(sec, "change_charges")) {  // qatom.f90:731-764
        auto tok = split_ws(ln);
        int iat = to_int(at(tok, 0, "change_charges"));
        if (!valid_q(iat)) throw Die(">>>>> ERROR: " + std::to_string(iat) + " is not a valid q-atom number");
        for (int s = 0; s < ns; s++) f.qcrg[(size_t)(iat - 1) * ns + s] = to_float(at(tok, 1 + s, "charge"));
    }

    const auto &types = section(sec, "atom_types");  // qatom.f90:780-823
    bool vdw_from_topo = false;
    std::map<std::string, int> index;
    if (!types.empty()) {
        f.nqlib = (int)types.size();
        f.qavdw.assign((size_t)f.nqlib * NLJTYP, 0.0);
        f.qbvdw.assign((size_t)f.nqlib * NLJTYP, 0.0);
        for (int i = 0; i < f.nqlib; i++) {
            auto tok = split_ws(types[i]);
            f.qtac.push_back(at(tok, 0, "atom type name"));
            for (int k = 0; k < NLJTYP; k++) {
                f.qavdw[(size_t)i * NLJTYP + k] = to_float(at(tok, 1 + 2 * k, "atom type A"));
                f.qbvdw[(size_t)i * NLJTYP + k] = to_float(at(tok, 2 + 2 * k, "atom type B"));
            }
            if (index.count(tok[0]))
                throw Die(">>>>> ERROR: Could not enumerate q-atom type " + tok[0] + " Duplicate name?");
            index[tok[0]] = i + 1;
        }
    } else {
        vdw_from_topo = true;
        f.nqlib = nq;
        f.qavdw.assign((size_t)nq * NLJTYP, 0.0);
        f.qbvdw.assign((size_t)nq * NLJTYP, 0.0);
        for (int i = 0; i < nq; i++) {
            int t = topo.iac[f.iqseq[i] - 1];
            f.qtac.push_back(!topo.tac.empty() ? topo.tac[t - 1] : std::to_string(t));
            for (int k = 0; k < NLJTYP; k++) {
                f.qavdw[(size_t)i * NLJTYP + k] = topo.iaclib[7 * (t - 1) + 1 + k];
                f.qbvdw[(size_t)i * NLJTYP + k] = topo.iaclib[7 * (t - 1) + 4 + k];
            }
        }
    }

    f.qiac.assign((size_t)nq * ns, 0);
    const auto &chg = section(sec, "change_atoms");  // qatom.f90:825-866
    if (chg.empty()) {
        f.qvdw_flag = false;
        if (vdw_from_topo)
            for (int i = 0; i < nq; i++)
                for (int s = 0; s < ns; s++) f.qiac[(size_t)i * ns + s] = i + 1;
    } else {
        int mx = 0;
        for (auto &ln : chg) mx = std::max(mx, to_int(at(split_ws(ln), 0, "change_atoms")));
        if (mx != nq || (int)chg.size() != nq)
            throw Die(">>>>> ERROR: Atom types of Q-atoms must be given for every Q-atom!");
        f.qvdw_flag = true;
        for (auto &ln : chg) {
            auto tok = split_ws(ln);
            int iat = to_int(tok[0]);
            if (iat < 1) throw Die(">>>>> ERROR: invalid q-atom number in [change_atoms]");
            for (int s = 0; s < ns; s++) {
                const std::string &nm = at(tok, 1 + s, "q-atom type");
                if (!index.count(nm)) throw Die(">>>>> ERROR: Q-atom type " + nm + " has not been defined.");
                f.qiac[(size_t)(iat - 1) * ns + s] = index[nm];
            }
        }
    }

    for (auto &ln : section(sec, "soft_pairs")) {  // qatom.f90:872-897
        auto tok = split_ws(ln);
        int j = to_int(at(tok, 0, "soft pair")), k = to_int(at(tok, 1, "soft pair"));
        if (!valid_q(j) || !valid_q(k))
            throw Die(">>>>> ERROR: Invalid q-atom number in this group: " + std::to_string(j) + " " + std::to_string(k));
        f.iqexpnb.push_back(j);
        f.jqexpnb.push_back(k);
    }

    for (auto &ln : section(sec, "el_scale")) {  // qatom.f90:903-930
        auto tok = split_ws(ln);
        int j = to_int(at(tok, 0, "el_scale")), k = to_int(at(tok, 1, "el_scale"));
        if (!valid_q(j) || !valid_q(k))
            throw Die(">>>>> ERROR: Invalid q-atom number in this group: " + std::to_string(j) + " " + std::to_string(k));
        f.el_scale_iq.push_back(j);
        f.el_scale_jq.push_back(k);
        for (int s = 0; s < ns; s++) f.el_scale.push_back(to_float(at(tok, 2 + s, "el_scale value")));
    }

    for (auto &ln : section(sec, "excluded_pairs")) {  // qatom.f90:936-965
        auto tok = split_ws(ln);
        f.exspec_ij.push_back(to_int(at(tok, 0, "excluded pair")) + offset);
        f.exspec_ij.push_back(to_int(at(tok, 1, "excluded pair")) + offset);
        for (int s = 0; s < ns; s++) {
            int v = to_int(at(tok, 2 + s, "excluded pair flag"));
            if (v != 0 && v != 1) throw Die(">>>>> ERROR: Special exclusion state flags are invalid.");
            f.exspec_flag.push_back(v);
        }
    }

    for (auto &ln : section(sec, "change_bonds")) {  // qatom.f90:990-1018
        auto tok = split_ws(ln);
        f.qbnd_ij.push_back(to_int(at(tok, 0, "change_bonds")) + offset);
        f.qbnd_ij.push_back(to_int(at(tok, 1, "change_bonds")) + offset);
        for (int s = 0; s < ns; s++) f.qbnd_cod.push_back(to_int(at(tok, 2 + s, "bond code")));
    }

    softcore(sec, f, topo);
}

}  // namespace

Fep qatom_load_fep(const std::string &path, const Topology &topo, int nstates_expected) {
    Sections sec = parse_sections(read_file(path, "fep"));
    Fep f;
    try {
        load_fep_impl(sec, topo, nstates_expected, f);
    } catch (const BadValue &e) {
        throw Die(std::string(">>>>> ERROR: Could not read fep file. (") + e.what() + ")");
    }
    return f;
}

// ------------------------------------------------------------------------------------------------ prep_sim
std::vector<int32_t> make_qconn(int nstates, int nat_solute, int nqat, const std::vector<int32_t> &iqseq,
                                const std::vector<int32_t> &iqatom, const std::vector<int32_t> &bnd_solute,
                                const std::vector<int32_t> &qbnd_ij, const std::vector<int32_t> &qbnd_cod,
                                const std::vector<int32_t> &exspec_ij, const std::vector<int32_t> &exspec_flag) {
    // qconn(state, atom, iq) = 1 for the Q-atom itself, 1 + number of bonds to it up to 4, 9 further away; walks
    // topology bonds with cod != 0 and the state's Q-bonds (qbnd%cod(state) > 0).  The reference's depth-first search
    // keeps the minimum level, i.e. the breadth-first distance computed here.  Special exclusions write 0.
    std::vector<int32_t> qconn((size_t)nqat * nat_solute * nstates, 9);
    std::vector<std::vector<int>> adj0(nat_solute + 1);
    for (size_t b = 0; b + 3 <= bnd_solute.size(); b += 3) {
        int i = bnd_solute[b], j = bnd_solute[b + 1], cod = bnd_solute[b + 2];
        if (cod == 0) continue;
        if (i >= 1 && j >= 1 && i <= nat_solute && j <= nat_solute) {
            adj0[i].push_back(j);
            adj0[j].push_back(i);
        }
    }
    const size_t nqb = qbnd_ij.size() / 2;
    std::vector<int> level(nat_solute + 1);
    for (int s = 0; s < nstates; s++) {
        auto adj = adj0;
        for (size_t b = 0; b < nqb; b++) {
            if (qbnd_cod[b * nstates + s] > 0) {
                int i = qbnd_ij[2 * b], j = qbnd_ij[2 * b + 1];
                if (i >= 1 && j >= 1 && i <= nat_solute && j <= nat_solute) {
                    adj[i].push_back(j);
                    adj[j].push_back(i);
                }
            }
        }
        for (int iq = 0; iq < nqat; iq++) {
            std::fill(level.begin(), level.end(), 0);
            int src = iqseq[iq];
            level[src] = 1;
            std::deque<int> dq{src};
            while (!dq.empty()) {
                int cur = dq.front();
                dq.pop_front();
                if (level[cur] >= 4) continue;
                for (int nb : adj[cur]) {
                    if (!level[nb]) {
                        level[nb] = level[cur] + 1;
                        dq.push_back(nb);
                    }
                }
            }
            for (int a = 1; a <= nat_solute; a++)
                if (level[a]) qconn[((size_t)iq * nat_solute + (a - 1)) * nstates + s] = level[a];
        }
    }
    const size_t nex = exspec_ij.size() / 2;
    for (size_t k = 0; k < nex; k++) {
        int pr[2][2] = {{exspec_ij[2 * k], exspec_ij[2 * k + 1]}, {exspec_ij[2 * k + 1], exspec_ij[2 * k]}};
        for (auto &p : pr) {
            int a = p[0], b = p[1];
            int iq = iqatom[a - 1];
            if (iq > 0 && b >= 1 && b <= nat_solute)
                for (int s = 0; s < nstates; s++)
                    if (exspec_flag[k * nstates + s]) qconn[((size_t)(iq - 1) * nat_solute + (b - 1)) * nstates + s] = 0;
        }
    }
    return qconn;
}

namespace {
template <class T>
const T *ptr_or_dummy(std::vector<T> &v) {
    if (v.empty()) v.push_back(T(0));  // the C ABI never reads past the stated sizes; keep pointers non-null
    return v.data();
}
}  // namespace

const qnb_system *System::view() {
    s.cgp = ptr_or_dummy(cgp);
    s.cgpatom = ptr_or_dummy(cgpatom);
    s.excl = ptr_or_dummy(excl);
    s.iqatom = ptr_or_dummy(iqatom);
    s.iqseq = ptr_or_dummy(iqseq);
    s.iac = ptr_or_dummy(iac);
    s.crg = ptr_or_dummy(crg);
    s.iaclib = ptr_or_dummy(iaclib);
    s.ljcod = ptr_or_dummy(ljcod);
    s.listex = ptr_or_dummy(listex);
    s.list14 = ptr_or_dummy(list14);
    s.listexlong = ptr_or_dummy(listexlong);
    s.list14long = ptr_or_dummy(list14long);
    s.qcrg = ptr_or_dummy(qcrg);
    s.qiac = ptr_or_dummy(qiac);
    s.qavdw = ptr_or_dummy(qavdw);
    s.qbvdw = ptr_or_dummy(qbvdw);
    s.sc_lookup = ptr_or_dummy(sc_lookup);
    s.iqexpnb = ptr_or_dummy(iqexpnb);
    s.jqexpnb = ptr_or_dummy(jqexpnb);
    s.el_scale_iq = ptr_or_dummy(el_scale_iq);
    s.el_scale_jq = ptr_or_dummy(el_scale_jq);
    s.el_scale = ptr_or_dummy(el_scale);
    s.qconn = ptr_or_dummy(qconn);
    return &s;
}

void System::full_shard() {
    s.pp_start = s.pw_start = s.qp_start = 1;
    s.pp_end = s.pw_end = s.qp_end = s.ncgp_solute;
    s.ww_start = s.qw_start = 1;
    s.ww_end = s.qw_end = s.nwat;
    s.natom_start = 1;
    s.natom_end = s.natom;
    s.is_master = 1;
}

std::vector<std::pair<int, int>> distribute_nonbonds(const std::vector<double> &per_i_counts, int nranks) {
    const int n = (int)per_i_counts.size();
    long long total = 0;
    for (double c : per_i_counts) total += (long long)c;
    std::vector<std::pair<int, int>> out;
    int i = 0;
    long long acc = 0;
    for (int r = 0; r < nranks; r++) {
        int start = i + 1;
        if (r == nranks - 1) {
            i = n;
        } else {
            long long target = (total * (r + 1)) / nranks;
            while (i < n && acc < target) {
                acc += (long long)per_i_counts[i];
                i++;
            }
        }
        out.push_back({start, i});
    }
    return out;
}

void System::shard(int rank, int nranks) {
    auto pr = distribute_nonbonds(std::vector<double>(s.ncgp_solute, 1.0), nranks)[rank];
    auto wr = distribute_nonbonds(std::vector<double>(s.nwat, 1.0), nranks)[rank];
    auto ar = distribute_nonbonds(std::vector<double>(s.natom, 1.0), nranks)[rank];
    s.pp_start = s.pw_start = s.qp_start = pr.first;
    s.pp_end = s.pw_end = s.qp_end = pr.second;
    s.ww_start = s.qw_start = wr.first;
    s.ww_end = s.qw_end = wr.second;
    s.natom_start = ar.first;
    s.natom_end = ar.second;
    s.is_master = rank == 0;
}

System prep_sim(const Topology &topo, const Fep *fep, bool use_LRF) {
    System q;
    qnb_system &s = q.s;
    s.abi_version = QNB_ABI_VERSION;
    s.natom = topo.nat_pro;
    s.nat_solute = topo.nat_solute;
    s.solv_atom = topo.solv_atom;
    s.nwat = topo.nwat;
    s.ncgp = topo.ncgp;
    s.ncgp_solute = topo.ncgp_solute;
    s.iuse_switch_atom = topo.iuse_switch_atom;
    s.natyps = topo.natyps;
    s.ivdw_rule = topo.ivdw_rule;
    s.el14_scale = topo.el14_scale;
    s.solvent_type = topo.solvent_type;
    s.use_PBC = topo.use_PBC;
    s.use_LRF = use_LRF;
    s.max_nbr_range = MAX_NBR_RANGE;
    s.rexcl_o = topo.rexcl_o;
    for (int k = 0; k < 3; k++) {
        s.xpcent[k] = topo.xpcent[k];
        q.boxlength[k] = topo.boxlength[k];
    }
    s.ntors_gt_solute = topo.ntors > topo.ntors_solute;
    q.cgp = topo.cgp;
    q.cgpatom = topo.cgpatom;
    q.iac = topo.iac;
    q.excl = topo.excl;
    q.xtop = topo.xtop;
    q.istart_mol = topo.istart_mol;
    // topology(): ljcod and sqrt(eps) (simprep.f90:4553-4571)
    s.num_atyp = topo.iac.empty() ? 0 : *std::max_element(topo.iac.begin(), topo.iac.end());
    q.ljcod.assign((size_t)s.num_atyp * s.num_atyp, 1);
    for (size_t k = 0; k + 1 < topo.lj2.size(); k += 2) {
        int i = topo.lj2[k], j = topo.lj2[k + 1];
        if (i < 1 || j < 1 || i > s.num_atyp || j > s.num_atyp) throw Die(">>>>> ERROR: LJ type 2 pair out of range");
        q.ljcod[(size_t)(i - 1) * s.num_atyp + (j - 1)] = 2;
        q.ljcod[(size_t)(j - 1) * s.num_atyp + (i - 1)] = 2;
    }
    q.iaclib = topo.iaclib;
    if (topo.ivdw_rule == 2)
        for (int i = 0; i < topo.natyps; i++)
            for (int k = 4; k < 7; k++) q.iaclib[7 * i + k] = std::sqrt(std::fabs(q.iaclib[7 * i + k]));
    q.mass.resize(s.natom);
    for (int i = 0; i < s.natom; i++) q.mass[i] = topo.iaclib[7 * (topo.iac[i] - 1)];
    q.listex = topo.listex;
    q.list14 = topo.list14;
    q.listexlong = topo.listexlong;
    q.list14long = topo.list14long;
    std::vector<double> crg = topo.crg;
    std::vector<int32_t> bnd(topo.bnd.begin(), topo.bnd.begin() + 3 * (size_t)std::min(topo.nbonds_solute, topo.nbonds));

    q.iqatom.assign(s.natom, 0);
    std::vector<double> qcrg;
    if (fep && fep->nqat > 0) {
        const Fep &f = *fep;
        s.nqat = f.nqat;
        s.nstates = f.nstates;
        s.qswitch = f.qswitch;
        const int nq = f.nqat, ns = f.nstates;
        q.iqseq = f.iqseq;
        for (int i = 0; i < nq; i++) {  // get_fep L1010-1019
            int a = f.iqseq[i];
            if (a > 0 && a <= s.nat_solute) q.iqatom[a - 1] = i + 1;
        }
        s.qvdw_flag = f.qvdw_flag;
        s.qq_use_library_charges = f.qq_use_library_charges;
        s.nqlib = f.nqlib;
        std::vector<double> qavdw = f.qavdw, qbvdw = f.qbvdw;  // [nqlib][3]
        if (f.qvdw_flag && topo.ivdw_rule == 2) {              // get_fep L1041-1046
            for (int i = 0; i < f.nqlib; i++) {
                qbvdw[(size_t)i * 3 + 0] = std::sqrt(qbvdw[(size_t)i * 3 + 0]);
                qbvdw[(size_t)i * 3 + 2] = std::sqrt(qbvdw[(size_t)i * 3 + 2]);
            }
        }
        // Fortran layouts: (type,code), (iq,state), (iq,k,state), (entry,state) column-major
        q.qavdw.assign((size_t)f.nqlib * 3, 0.0);
        q.qbvdw.assign((size_t)f.nqlib * 3, 0.0);
        for (int i = 0; i < f.nqlib; i++)
            for (int k = 0; k < 3; k++) {
                q.qavdw[(size_t)k * f.nqlib + i] = qavdw[(size_t)i * 3 + k];
                q.qbvdw[(size_t)k * f.nqlib + i] = qbvdw[(size_t)i * 3 + k];
            }
        q.qiac.assign((size_t)nq * ns, 0);
        for (int i = 0; i < nq; i++)
            for (int st = 0; st < ns; st++) q.qiac[(size_t)st * nq + i] = f.qiac[(size_t)i * ns + st];
        const int nk = topo.natyps + nq;
        q.sc_lookup.assign((size_t)nq * nk * ns, 0.0);
        for (int i = 0; i < nq; i++)
            for (int k = 0; k < nk; k++)
                for (int st = 0; st < ns; st++)
                    q.sc_lookup[((size_t)st * nk + k) * nq + i] = f.sc_lookup[((size_t)i * nk + k) * ns + st];
        q.iqexpnb = f.iqexpnb;
        q.jqexpnb = f.jqexpnb;
        s.nqexpnb = (int)f.iqexpnb.size();
        q.el_scale_iq = f.el_scale_iq;
        q.el_scale_jq = f.el_scale_jq;
        const int ne = (int)f.el_scale_iq.size();
        s.nel_scale = ne;
        q.el_scale.assign((size_t)ne * ns, 0.0);
        for (int e = 0; e < ne; e++)
            for (int st = 0; st < ns; st++) q.el_scale[(size_t)st * ne + e] = f.el_scale[(size_t)e * ns + st];
        // redefined bonds are switched off in the topology (get_fep L1049-1065)
        const size_t nqb = f.qbnd_ij.size() / 2;
        for (size_t b = 0; b + 3 <= bnd.size(); b += 3)
            for (size_t k = 0; k < nqb; k++) {
                int qi = f.qbnd_ij[2 * k], qj = f.qbnd_ij[2 * k + 1];
                if ((bnd[b] == qi && bnd[b + 1] == qj) || (bnd[b] == qj && bnd[b + 1] == qi)) bnd[b + 2] = 0;
            }
        // special exclusions involving a non-Q atom go into the exclusion lists (get_fep L1143-1179)
        const size_t nex = f.exspec_ij.size() / 2;
        for (size_t k = 0; k < nex; k++) {
            int i = f.exspec_ij[2 * k], j = f.exspec_ij[2 * k + 1];
            if (i < 1 || i > s.natom || j < 1 || j > s.natom) throw Die("invalid special exclusion data");
            if (q.iqatom[i - 1] == 0 || q.iqatom[j - 1] == 0) {
                bool any = false, all = true;
                for (int st = 0; st < ns; st++) {
                    bool on = f.exspec_flag[k * ns + st] != 0;
                    any |= on;
                    all &= on;
                }
                if (any) {
                    if (!all) throw Die("Non-Q-atom special excl. pair must be on in all or no states");
                    if (std::abs(j - i) <= MAX_NBR_RANGE) {
                        if (i < j)
                            q.listex[(size_t)(i - 1) * MAX_NBR_RANGE + (j - i - 1)] = 1;
                        else
                            q.listex[(size_t)(j - 1) * MAX_NBR_RANGE + (i - j - 1)] = 1;
                    } else {
                        q.listexlong.push_back(i);
                        q.listexlong.push_back(j);
                    }
                }
            }
        }
        q.qconn = make_qconn(ns, s.nat_solute, nq, q.iqseq, q.iqatom, bnd, f.qbnd_ij, f.qbnd_cod, f.exspec_ij,
                             f.exspec_flag);
        qcrg.assign((size_t)nq * ns, 0.0);
        for (int i = 0; i < nq; i++)
            for (int st = 0; st < ns; st++) qcrg[(size_t)st * nq + i] = f.qcrg[(size_t)i * ns + st];
    } else {
        s.nqat = 0;
        s.nstates = fep ? fep->nstates : 1;
    }
    s.nexlong = (int)q.listexlong.size() / 2;
    s.n14long = (int)q.list14long.size() / 2;
    // prep_sim: charges scaled by sqrt(coulomb_constant) (simprep.f90:3714-3721)
    const double sq = std::sqrt(topo.coulomb_constant);
    q.crg.resize(crg.size());
    for (size_t i = 0; i < crg.size(); i++) q.crg[i] = crg[i] * sq;
    q.qcrg.resize(qcrg.size());
    for (size_t i = 0; i < qcrg.size(); i++) q.qcrg[i] = qcrg[i] * sq;

    // init_constraints (simprep.f90:2167-2345) with the defaults of md.f90:338-357: shake_solvent on, shake_solute off,
    // shake_hydrogens on, shake_heavy off -> solvent bonds that involve a hydrogen (mass < 4, topo.f90:896-903)
    for (size_t b = 0; b + 3 <= topo.bnd.size(); b += 3) {
        int ia = topo.bnd[b], ja = topo.bnd[b + 1], cod = topo.bnd[b + 2];
        if (cod == 0 || ia <= s.nat_solute) continue;
        bool has_h = q.mass[ia - 1] < 4.0 || q.mass[ja - 1] < 4.0;
        if (!has_h) continue;
        if (cod < 1 || 2 * (size_t)cod > topo.bondlib.size()) throw Die(">>>>> ERROR: bond code without library entry");
        double b0 = topo.bondlib[2 * (cod - 1) + 1];
        q.const_ij.push_back(ia);
        q.const_ij.push_back(ja);
        q.const_dist2.push_back(b0 * b0);
    }
    q.full_shard();
    q.view();
    return q;
}

System system_from_struct(const qnb_system &src, const double boxlength[3]) {
    if (src.abi_version != QNB_ABI_VERSION) throw Die("system_from_struct: ABI version mismatch");
    System q;
    q.s = src;
    for (int k = 0; k < 3; k++) q.boxlength[k] = boxlength ? boxlength[k] : 0.0;
    auto ci = [](const int32_t *p, size_t n) { return p && n ? std::vector<int32_t>(p, p + n) : std::vector<int32_t>(); };
    auto cd = [](const double *p, size_t n) { return p && n ? std::vector<double>(p, p + n) : std::vector<double>(); };
    const size_t nat = src.natom, nsol = src.nat_solute, nq = src.nqat, ns = src.nstates;
    q.cgp = ci(src.cgp, 3 * (size_t)src.ncgp);
    size_t ncgpatom = 0;  // cgp(ncgp)%last: not every topology puts all atoms into charge groups
    for (size_t g = 0; g < (size_t)src.ncgp; g++) ncgpatom = std::max(ncgpatom, (size_t)std::max(q.cgp[3 * g + 2], 0));
    q.cgpatom = ci(src.cgpatom, ncgpatom);
    q.excl = ci(src.excl, nat);
    q.iqatom = ci(src.iqatom, nat);
    q.iqseq = ci(src.iqseq, nq);
    q.iac = ci(src.iac, nat);
    q.crg = cd(src.crg, nat);
    q.iaclib = cd(src.iaclib, 7 * (size_t)src.natyps);
    q.ljcod = ci(src.ljcod, (size_t)src.num_atyp * src.num_atyp);
    q.listex = ci(src.listex, (size_t)src.max_nbr_range * nsol);
    q.list14 = ci(src.list14, (size_t)src.max_nbr_range * nsol);
    q.listexlong = ci(src.listexlong, 2 * (size_t)src.nexlong);
    q.list14long = ci(src.list14long, 2 * (size_t)src.n14long);
    q.qcrg = cd(src.qcrg, nq * ns);
    q.qiac = ci(src.qiac, nq * ns);
    q.qavdw = cd(src.qavdw, 3 * (size_t)src.nqlib);
    q.qbvdw = cd(src.qbvdw, 3 * (size_t)src.nqlib);
    q.sc_lookup = cd(src.sc_lookup, nq * ((size_t)src.natyps + nq) * ns);
    q.iqexpnb = ci(src.iqexpnb, src.nqexpnb);
    q.jqexpnb = ci(src.jqexpnb, src.nqexpnb);
    q.el_scale_iq = ci(src.el_scale_iq, src.nel_scale);
    q.el_scale_jq = ci(src.el_scale_jq, src.nel_scale);
    q.el_scale = cd(src.el_scale, (size_t)src.nel_scale * ns);
    q.qconn = ci(src.qconn, ns * nsol * nq);
    q.mass.resize(nat);
    for (size_t i = 0; i < nat; i++) q.mass[i] = q.iaclib[7 * (size_t)(q.iac[i] - 1)];
    q.view();
    return q;
}

// ------------------------------------------------------------------------------------------------ SHAKE
namespace {
constexpr double CONST_TOL = 0.0001;  // globals.f90:519
constexpr int CONST_MAX_ITER = 1000;  // globals.f90:520
}  // namespace

std::vector<int32_t> constraint_molecules(const System &sys) {
    std::vector<int32_t> first{0};
    const size_t nc = sys.const_dist2.size();
    auto mol_of = [&](int atom) {
        auto it = std::upper_bound(sys.istart_mol.begin(), sys.istart_mol.end(), atom);
        return (int)(it - sys.istart_mol.begin()) - 1;
    };
    for (size_t c = 1; c <= nc; c++)
        if (c == nc || mol_of(sys.const_ij[2 * c]) != mol_of(sys.const_ij[2 * (c - 1)])) first.push_back((int32_t)c);
    return first;
}

int shake(const System &sys, const double *xx, double *x) {
    // bondene.f90:1069-1150, literal order of operations: a constraint found within tolerance is flagged ready and
    // still receives that pass's correction; ready constraints are not looked at again in this call.
    const int natom = sys.s.natom;
    const std::vector<int32_t> first = constraint_molecules(sys);
    long total = 0;
    std::vector<char> ready;
    for (size_t m = 0; m + 1 < first.size(); m++) {
        const size_t c0 = first[m], c1 = first[m + 1];
        ready.assign(c1 - c0, 0);
        int nits = 0;
        for (;;) {
            for (size_t ic = c0; ic < c1; ic++) {
                if (ready[ic - c0]) continue;
                const int i = sys.const_ij[2 * ic] - 1, j = sys.const_ij[2 * ic + 1] - 1;
                if (i < 0 || j < 0 || i >= natom || j >= natom) throw Die("shake: constraint atom out of range");
                const double dist2 = sys.const_dist2[ic];
                const double xij[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
                const double diff = dist2 - (xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2]);
                if (std::fabs(diff) < CONST_TOL * dist2) ready[ic - c0] = 1;
                const double xxij[3] = {xx[3 * i] - xx[3 * j], xx[3 * i + 1] - xx[3 * j + 1],
                                        xx[3 * i + 2] - xx[3 * j + 2]};
                const double scp = xij[0] * xxij[0] + xij[1] * xxij[1] + xij[2] * xxij[2];
                const double wi = 1.0 / sys.mass[i], wj = 1.0 / sys.mass[j];
                const double corr = diff / (2.0 * scp * (wi + wj));
                for (int k = 0; k < 3; k++) {
                    x[3 * i + k] = x[3 * i + k] + xxij[k] * corr * wi;
                    x[3 * j + k] = x[3 * j + k] + (-xxij[k]) * corr * wj;
                }
            }
            nits++;
            bool all = true;
            for (char r : ready) all = all && r;
            if (all) break;
            if (nits >= CONST_MAX_ITER) throw Die("shake failure");  // bondene.f90:1143
        }
        total += nits;
    }
    const int nmol = (int)sys.istart_mol.size();
    return nmol > 0 ? (int)(total / nmol) : 0;
}

int initial_constraint(const System &sys, double *x) {
    std::vector<double> xx(x, x + 3 * (size_t)sys.s.natom);
    return shake(sys, xx.data(), x);
}

// ------------------------------------------------------------------------------------------------ Nonbonded
Nonbonded::Nonbonded(System &sys, int device) : sys_(sys) {
    const qnb_system *s = sys_.view();
    x.assign(3 * (size_t)s->natom, 0.0);
    d.assign(3 * (size_t)s->natom, 0.0);
    if (qnb_init(s, device, &h_) != 0) throw Die(std::string("qnb_init: ") + qnb_last_error());
    // x and d live as long as this object (md.f90's module arrays): page-locked once, so that every step uploads x without
    // a staging copy and the device adds the gradient into d
    if (qnb_register_host_buffers(h_, x.data(), d.data()) != 0) throw Die(std::string("qnb_register_host_buffers: ") + qnb_last_error());
    if (s->use_PBC) update_box(sys_.boxlength);
}

Nonbonded::~Nonbonded() {
    if (h_) qnb_finalize(h_);
}

void Nonbonded::update_box(const double boxlength[3]) {
    double inv[3];
    for (int k = 0; k < 3; k++) {
        sys_.boxlength[k] = boxlength[k];
        inv[k] = 1.0 / boxlength[k];
    }
    if (qnb_update_box(h_, sys_.boxlength, inv) != 0) throw Die(std::string("qnb_update_box: ") + qnb_last_error());
}

void Nonbonded::save_lists() {
    if (qnb_save_lists(h_) != 0) throw Die(std::string("MC_volume (save lists): ") + qnb_last_error());
}

void Nonbonded::restore_lists() {
    if (qnb_restore_lists(h_) != 0) throw Die(std::string("MC_volume (restore lists): ") + qnb_last_error());
}

void Nonbonded::set_solvent_restraints(const qnb_solvent_restraints &p) {
    if (qnb_set_solvent_restraints(h_, &p) != 0) throw Die(std::string("restrain_solvent: ") + qnb_last_error());
}

void Nonbonded::set_theta_corr(const double *theta_corr) {
    if (qnb_set_theta_corr(h_, theta_corr) != 0) throw Die(std::string("watpol: ") + qnb_last_error());
}

void Nonbonded::last_restraints(double E[2], double *shell_theta_sum, int32_t *shell_n) {
    if (qnb_last_restraints(h_, E, shell_theta_sum, shell_n) != 0) throw Die(std::string("watpol: ") + qnb_last_error());
}

void Nonbonded::set_constraints() {
    const std::vector<int32_t> first = constraint_molecules(sys_);
    const size_t nc = sys_.const_dist2.size();
    std::vector<double> winv(sys_.mass.size());
    for (size_t i = 0; i < winv.size(); i++) winv[i] = 1.0 / sys_.mass[i];   // simprep.f90:3675
    static const int32_t no_ij[2] = {0, 0};
    static const double no_d2[1] = {0.0};
    if (qnb_set_constraints(h_, (int)first.size() - 1, first.data(), nc ? sys_.const_ij.data() : no_ij,
                            nc ? sys_.const_dist2.data() : no_d2, winv.data()) != 0)
        throw Die(std::string("init_constraints: ") + qnb_last_error());
}

int64_t Nonbonded::shake(const double *xx, double *x) {
    int64_t it = 0;
    if (qnb_shake(h_, xx, x, &it) != 0) throw Die(qnb_last_error());   // 'shake failure', bondene.f90:1143
    return it;
}

void Nonbonded::qcp_beads(const double *x_save, int natq, const int32_t *atoms, int nbeads, const double *coord,
                          const double *lambda, double *EQ_out) {
    if (qnb_qcp_beads(h_, x_save, natq, atoms, nbeads, coord, lambda, EQ_out) != 0)
        throw Die(std::string("qcp_run: ") + qnb_last_error());
}

void Nonbonded::make_pair_lists(double Rq, double Rcq2, double RcLRF2, double Rcpp2, double Rcpw2, double Rcww2) {
    if (qnb_build_lists(h_, x.data(), Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF, dump ? nb_pairs : nullptr) != 0)
        throw Die(std::string("make_pair_lists: ") + qnb_last_error());
}

void Nonbonded::pot_energy_nonbonds(ENERGIES &E_loc, std::vector<OQ_ENERGIES> &EQ_loc, bool md) {
    const int ns = sys_.s.nstates;
    if ((int)EQ_loc.size() != ns) throw Die("pot_energy_nonbonds: EQ must have nstates entries");
    std::vector<double> lambda(ns), EQ(QNB_EQ_STRIDE * (size_t)ns, 0.0);
    for (int s = 0; s < ns; s++) lambda[s] = EQ_loc[s].lambda;
    double E[QNB_E_COUNT] = {0};
    const int flags = (md ? QNB_FLAG_MD : 0) | QNB_FLAG_QQ | ((md && restraints_on) ? QNB_FLAG_SOLVENT_RESTRAINTS : 0);
    if (qnb_nonbond(h_, x.data(), lambda.data(), flags, d.data(), E, EQ.data()) != 0)
        throw Die(std::string("pot_energy_nonbonds: ") + qnb_last_error());
    E_loc.pp = {E[QNB_E_PP_EL], E[QNB_E_PP_VDW]};
    E_loc.pw = {E[QNB_E_PW_EL], E[QNB_E_PW_VDW]};
    E_loc.ww = {E[QNB_E_WW_EL], E[QNB_E_WW_VDW]};
    E_loc.LRF = E[QNB_E_LRF];
    for (int s = 0; s < ns; s++) {
        const double *e = &EQ[QNB_EQ_STRIDE * (size_t)s];
        EQ_loc[s].qq = {e[0], e[1]};
        EQ_loc[s].qp = {e[2], e[3]};
        EQ_loc[s].qw = {e[4], e[5]};
    }
}

// ------------------------------------------------------------------------------------------------ write_out
std::string write_out_nonbonded(const System &sys, const ENERGIES &E, const std::vector<OQ_ENERGIES> &EQ, int istep) {
    // qalloc.f90:686-806; formats 2, 4, 6 (A,T17,6F12.2), 26, 32 (a,T8,i2,f7.4,2f10.2)
    std::string out;
    char buf[256];
    auto row6 = [&](const char *name, const double *v, int n) {
        std::string ln = name;
        if (ln.size() < 16) ln.resize(16, ' ');
        for (int k = 0; k < n; k++) {
            std::snprintf(buf, sizeof buf, "%12.2f", v[k]);
            ln += buf;
        }
        out += ln + "\n";
    };
    auto row32 = [&](const char *name, int istate, double lambda, double a, double b) {
        std::string ln = name;
        if (ln.size() < 7) ln.resize(7, ' ');
        std::snprintf(buf, sizeof buf, "%2d%7.4f%10.2f%10.2f", istate, lambda, a, b);
        out += ln + buf + "\n";
    };
    std::snprintf(buf, sizeof buf, "======================= %-15s at step %6d ========================\n", "Energy summary",
                  istep);
    out += buf;
    std::snprintf(buf, sizeof buf, "%16s%12s%12s\n", "", "el", "vdW");
    out += buf;
    const double pp[2] = {E.pp.el, E.pp.vdw}, ww[2] = {E.ww.el, E.ww.vdw}, pw[2] = {E.pw.el, E.pw.vdw};
    row6("solute", pp, 2);
    if (sys.s.nwat > 0) row6("solvent", ww, 2);
    row6("solute-solvent", pw, 2);
    if (sys.s.use_LRF) row6("LRF", &E.LRF, 1);
    if (sys.s.nstates > 0 && sys.s.nqat > 0) {
        std::snprintf(buf, sizeof buf, "======================= %-15s at step %6d ========================\n",
                      "Q-atom energies", istep);
        out += buf;
        std::snprintf(buf, sizeof buf, "type   st lambda%10s%10s\n", "el", "vdW");
        out += buf;
        const int ns = (int)EQ.size();
        for (int s = 0; s < ns; s++) row32("Q-Q", s + 1, EQ[s].lambda, EQ[s].qq.el, EQ[s].qq.vdw);
        out += "\n";
        if (sys.s.nat_solute > sys.s.nqat) {
            for (int s = 0; s < ns; s++) row32("Q-prot", s + 1, EQ[s].lambda, EQ[s].qp.el, EQ[s].qp.vdw);
            out += "\n";
        }
        if (sys.s.nwat > 0) {
            for (int s = 0; s < ns; s++) row32("Q-wat", s + 1, EQ[s].lambda, EQ[s].qw.el, EQ[s].qw.vdw);
            out += "\n";
        }
        for (int s = 0; s < ns; s++)
            row32("Q-surr.", s + 1, EQ[s].lambda, EQ[s].qp.el + EQ[s].qw.el, EQ[s].qp.vdw + EQ[s].qw.vdw);
        out += "\n";
    }
    return out;
}

}  // namespace qdyn
