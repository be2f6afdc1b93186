// Host-side model reader and propensity compiler.
//
// Behavioural contract (what has to match the reference, not how it is written):
//  * `.input` grammar: keyword line then values; sections nspecies nreactions nparameters
//    species parameters reactions propensities; unknown lines are skipped
//    (src/model/ModelModule.f90:91-158).  Keywords are matched case-insensitively, a
//    documented fix of the reference's upper-case literals.
//  * reaction strings: blank-separated terms, "->" or "<-", "+" ignored, "0" = nothing,
//    optional integer prefix = coefficient, species found by substring search with the
//    stale-coefficient behaviour of ModelModule.f90:272-290.
//  * propensity strings: '**' == '^', blanks ignored; the expression is split at the
//    RIGHTMOST top-level binary operator, trying + - * / ^ in that order
//    (src/parser/FortranParser.f90:679-706) -- so a*b/c means a*(b/c) and a^b^c means
//    (a^b)^c -- and '^' is always pow(double,double).  That order decides last-bit values.
//  * evaluation: division by zero, log of a non-positive number, sqrt of a negative and
//    asin/acos out of range give 0 (FortranParser.f90:217-287).
#include "model_host.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

namespace kfsp {

namespace {

const char* kFuncNames[] = {"abs", "exp", "log10", "log", "sqrt", "sinh", "cosh", "tanh",
                            "sin", "cos", "tan", "asin", "acos", "atan"};
const int kNumFuncs = 14;

struct Node {
    int32_t op = 0;                 // cImmed, cNeg, binary op, function, or >= VarBegin
    double value = 0.0;
    std::unique_ptr<Node> a, b;
};

struct Compiler {
    std::string f;                  // condensed expression
    const std::vector<std::string>& vars;
    std::string err;

    Compiler(const std::string& s, const std::vector<std::string>& v) : f(s), vars(v) {}

    static bool is_alpha(char c) { return std::isalpha((unsigned char)c) != 0; }
    static bool is_digit(char c) { return c >= '0' && c <= '9'; }
    static bool in_set(char c, const char* set) { return c != '\0' && std::strchr(set, c) != nullptr; }

    // index of the function whose name prefixes f[b..e] (case-insensitive), first match in table order
    int func_at(int b, int e) const {
        for (int k = 0; k < kNumFuncs; ++k) {
            int len = (int)std::strlen(kFuncNames[k]);
            int have = e - b + 1;
            int cmp = len < have ? len : have;
            bool ok = cmp == len;             // a shorter remainder cannot equal the padded name
            for (int i = 0; ok && i < cmp; ++i)
                if (std::tolower((unsigned char)f[b + i]) != kFuncNames[k][i]) ok = false;
            if (ok) return k;
        }
        return -1;
    }
    bool enclosed(int b, int e) const {
        if (b > e || b < 0 || f[b] != '(' || f[e] != ')') return false;
        int depth = 0;
        for (int j = b + 1; j < e; ++j) {
            if (f[j] == '(') ++depth;
            else if (f[j] == ')') --depth;
            if (depth < 0) break;
        }
        return depth == 0;
    }
    bool binary_at(int j) const {
        if (f[j] != '+' && f[j] != '-') return true;
        if (j == 0) return false;
        if (in_set(f[j - 1], "+-*/^(")) return false;
        if (j + 1 < (int)f.size() && is_digit(f[j + 1]) && in_set(f[j - 1], "eEdD")) {
            bool digit = false, point = false;
            int k = j - 1;
            while (k > 0) {
                --k;
                if (is_digit(f[k])) digit = true;
                else if (f[k] == '.') { if (point) break; point = true; }
                else break;
            }
            if (digit && (k == 0 || in_set(f[k], "+-*/^("))) return false;
        }
        return true;
    }
    bool number(int b, int e, double& out) {
        // [digits][.digits][(e|E|d|D)[+|-]digits]
        int i = b;
        bool man = false, ex = false;
        std::string txt;
        while (i <= e && is_digit(f[i])) { txt += f[i++]; man = true; }
        if (i <= e && f[i] == '.') { txt += f[i++]; while (i <= e && is_digit(f[i])) { txt += f[i++]; man = true; } }
        if (i <= e && in_set(f[i], "eEdD")) {
            txt += 'e'; ++i;
            if (i <= e && (f[i] == '+' || f[i] == '-')) txt += f[i++];
            while (i <= e && is_digit(f[i])) { txt += f[i++]; ex = true; }
            if (!ex) return false;
        }
        if (!man || i != e + 1) return false;
        out = std::strtod(txt.c_str(), nullptr);
        return true;
    }
    std::unique_ptr<Node> leaf(int b, int e) {
        auto n = std::make_unique<Node>();
        if (b > e) { err = "missing operand"; return nullptr; }
        if (is_digit(f[b]) || f[b] == '.') {
            if (!number(b, e, n->value)) { err = "invalid number format: " + f.substr(b, e - b + 1); return nullptr; }
            n->op = cImmed;
            return n;
        }
        int stop = b;
        while (stop <= e && !in_set(f[stop], "+-*/^) ")) ++stop;
        std::string name = f.substr(b, stop - b);
        for (size_t j = 0; j < vars.size(); ++j)
            if (vars[j] == name) { n->op = VarBegin + (int32_t)j; return n; }
        err = "invalid element: " + f.substr(b, e - b + 1);
        return nullptr;
    }
    std::unique_ptr<Node> unary(int32_t op, std::unique_ptr<Node> x) {
        if (!x) return nullptr;
        auto n = std::make_unique<Node>();
        n->op = op;
        n->a = std::move(x);
        return n;
    }
    std::unique_ptr<Node> build(int b, int e) {
        if (b > e) { err = "missing operand"; return nullptr; }
        if (f[b] == '+') return build(b + 1, e);
        if (enclosed(b, e)) return build(b + 1, e - 1);
        if (is_alpha(f[b])) {
            int k = func_at(b, e);
            if (k >= 0) {
                size_t p = f.find('(', b);
                if (p != std::string::npos && (int)p <= e && enclosed((int)p, e))
                    return unary(cAbs + k, build((int)p + 1, e - 1));
            }
        } else if (f[b] == '-') {
            if (enclosed(b + 1, e)) return unary(cNeg, build(b + 2, e - 1));
            if (b + 1 <= e && is_alpha(f[b + 1])) {
                int k = func_at(b + 1, e);
                if (k >= 0) {
                    size_t p = f.find('(', b + 1);
                    if (p != std::string::npos && (int)p <= e && enclosed((int)p, e))
                        return unary(cNeg, unary(cAbs + k, build((int)p + 1, e - 1)));
                }
            }
        }
        static const char ops[] = {'+', '-', '*', '/', '^'};
        for (int io = 0; io < 5; ++io) {
            int depth = 0;
            for (int j = e; j >= b; --j) {
                if (f[j] == ')') ++depth;
                else if (f[j] == '(') --depth;
                if (depth == 0 && f[j] == ops[io] && binary_at(j)) {
                    if (io >= 2 && f[b] == '-') return unary(cNeg, build(b + 1, e));
                    auto n = std::make_unique<Node>();
                    n->op = cAdd + io;
                    n->a = build(b, j - 1);
                    n->b = build(j + 1, e);
                    if (!n->a || !n->b) return nullptr;
                    return n;
                }
            }
        }
        if (f[b] == '-') return unary(cNeg, leaf(b + 1, e));
        return leaf(b, e);
    }
    static void emit(const Node* n, Program& out) {
        if (n->a) emit(n->a.get(), out);
        if (n->b) emit(n->b.get(), out);
        out.code.push_back(n->op);
        if (n->op == cImmed) out.immed.push_back(n->value);
    }
};

std::string upper(std::string s) {
    for (auto& c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}
std::string first_token(const std::string& ln) {
    std::istringstream is(ln);
    std::string t;
    is >> t;
    return t;
}
bool blank(const std::string& s) {
    for (char c : s) if (!std::isspace((unsigned char)c)) return false;
    return true;
}

}  // namespace

int program_stack_depth(const Program& p, int nvars) {
    int sp = 0, mx = 0;
    size_t imm = 0;
    for (int32_t op : p.code) {
        if (op == cImmed) { if (imm++ >= p.immed.size()) return -1; ++sp; }
        else if (op == cNeg || (op >= cAbs && op <= cAtan)) { if (sp < 1) return -1; }
        else if (op >= cAdd && op <= cPow) { if (sp < 2) return -1; --sp; }
        else if (op >= VarBegin && op < VarBegin + nvars) ++sp;
        else return -1;
        if (sp > mx) mx = sp;
    }
    return sp == 1 ? mx : -1;
}

void program_profile(const Program& p, int S, uint32_t* species_mask, bool* inexact) {
    uint32_t mask = 0;
    bool ix = false;
    for (int32_t op : p.code) {
        if (op >= VarBegin && op < VarBegin + S) mask |= 1u << (op - VarBegin);
        else if (op == cPow || op == cExp || op == cLog10 || op == cLog || (op >= cSinh && op <= cAtan)) ix = true;
    }
    *species_mask = mask;
    *inexact = ix;
}

bool compile_expression(const std::string& expr, const std::vector<std::string>& vars, Program& out, std::string& err) {
    std::string s;
    for (size_t i = 0; i < expr.size(); ++i) {
        if (expr[i] == '*' && i + 1 < expr.size() && expr[i + 1] == '*') { s += '^'; ++i; continue; }
        if (std::isspace((unsigned char)expr[i])) continue;
        s += expr[i];
    }
    if (s.empty()) { err = "empty propensity expression"; return false; }
    int depth = 0;
    for (char c : s) {
        if (c == '(') ++depth;
        if (c == ')') --depth;
        if (depth < 0) { err = "mismatched parenthesis in: " + expr; return false; }
    }
    if (depth != 0) { err = "missing ) in: " + expr; return false; }
    Compiler c(s, vars);
    auto root = c.build(0, (int)s.size() - 1);
    if (!root) { err = c.err + " in: " + expr; return false; }
    out = Program();
    Compiler::emit(root.get(), out);
    out.stack_depth = program_stack_depth(out, (int)vars.size());
    if (out.stack_depth < 0) { err = "malformed expression: " + expr; return false; }
    return true;
}

double evaluate_program_checked(const Program& p, const double* val, bool* aborted) {
    double st[64];
    int sp = -1;
    size_t dp = 0;
    if (aborted) *aborted = false;
#define KFSP_ABORT() do { if (aborted) *aborted = true; return 0.0; } while (0)
    for (int32_t op : p.code) {
        switch (op) {
        case cImmed: st[++sp] = p.immed[dp++]; break;
        case cNeg: st[sp] = -st[sp]; break;
        case cAdd: st[sp - 1] = st[sp - 1] + st[sp]; --sp; break;
        case cSub: st[sp - 1] = st[sp - 1] - st[sp]; --sp; break;
        case cMul: st[sp - 1] = st[sp - 1] * st[sp]; --sp; break;
        case cDiv: if (st[sp] == 0.0) KFSP_ABORT(); st[sp - 1] = st[sp - 1] / st[sp]; --sp; break;
        case cPow: st[sp - 1] = std::pow(st[sp - 1], st[sp]); --sp; break;
        case cAbs: st[sp] = std::fabs(st[sp]); break;
        case cExp: st[sp] = std::exp(st[sp]); break;
        case cLog10: if (st[sp] <= 0.0) KFSP_ABORT(); st[sp] = std::log10(st[sp]); break;
        case cLog: if (st[sp] <= 0.0) KFSP_ABORT(); st[sp] = std::log(st[sp]); break;
        case cSqrt: if (st[sp] < 0.0) KFSP_ABORT(); st[sp] = std::sqrt(st[sp]); break;
        case cSinh: st[sp] = std::sinh(st[sp]); break;
        case cCosh: st[sp] = std::cosh(st[sp]); break;
        case cTanh: st[sp] = std::tanh(st[sp]); break;
        case cSin: st[sp] = std::sin(st[sp]); break;
        case cCos: st[sp] = std::cos(st[sp]); break;
        case cTan: st[sp] = std::tan(st[sp]); break;
        case cAsin: if (st[sp] < -1.0 || st[sp] > 1.0) KFSP_ABORT(); st[sp] = std::asin(st[sp]); break;
        case cAcos: if (st[sp] < -1.0 || st[sp] > 1.0) KFSP_ABORT(); st[sp] = std::acos(st[sp]); break;
        case cAtan: st[sp] = std::atan(st[sp]); break;
        default: st[++sp] = val[op - VarBegin]; break;
        }
    }
#undef KFSP_ABORT
    return st[0];
}
double evaluate_program(const Program& p, const double* val) { return evaluate_program_checked(p, val, nullptr); }

bool factor_program(const Program& p, int S, Factored& out) {
    // postfix -> tree: a sub-expression is a contiguous range of the code (and of the immediates)
    struct TNode { int32_t op; int a, b; int c0, c1, i0, i1; uint32_t mask; };
    std::vector<TNode> nodes;
    std::vector<int> stack;
    int imm = 0;
    for (int ip = 0; ip < (int)p.code.size(); ++ip) {
        const int32_t op = p.code[ip];
        TNode n{op, -1, -1, ip, ip + 1, imm, imm, 0u};
        if (op == cImmed) {
            if (imm >= (int)p.immed.size()) return false;
            n.i1 = ++imm;
        } else if (op == cNeg || (op >= cAbs && op <= cAtan)) {
            if (stack.empty()) return false;
            n.a = stack.back(); stack.pop_back();
        } else if (op >= cAdd && op <= cPow) {
            if (stack.size() < 2) return false;
            n.b = stack.back(); stack.pop_back();
            n.a = stack.back(); stack.pop_back();
        } else if (op >= VarBegin) {
            if (op - VarBegin < S) n.mask = 1u << (op - VarBegin);
        } else {
            return false;
        }
        if (n.a >= 0) { n.c0 = nodes[n.a].c0; n.i0 = nodes[n.a].i0; n.mask |= nodes[n.a].mask; }
        if (n.b >= 0) n.mask |= nodes[n.b].mask;
        n.i1 = imm;
        nodes.push_back(n);
        stack.push_back((int)nodes.size() - 1);
    }
    if (stack.size() != 1) return false;
    out = Factored();
    bool ok = true;
    auto single = [](uint32_t m) { return (m & (m - 1)) == 0; };
    auto reduce = [&](auto&& self, int id) -> void {
        const TNode& n = nodes[id];
        if (single(n.mask)) {
            FactoredTerm t;
            uint32_t m = n.mask;
            while (m > 1) { m >>= 1; ++t.species; }
            if (n.mask) ++t.species;
            t.prog.code.assign(p.code.begin() + n.c0, p.code.begin() + n.c1);
            t.prog.immed.assign(p.immed.begin() + n.i0, p.immed.begin() + n.i1);
            out.ops.push_back((int32_t)out.terms.size());
            out.terms.push_back(t);
            return;
        }
        if (n.op == cNeg) { self(self, n.a); out.ops.push_back(-cNeg); return; }
        if (n.op == cAdd || n.op == cSub || n.op == cMul) {
            self(self, n.a);
            self(self, n.b);
            out.ops.push_back(-n.op);
            return;
        }
        ok = false;
    };
    reduce(reduce, stack.back());
    return ok;
}

double HostModel::propensity(const int32_t* state, int reaction1) const {
    if (custom) return custom(state, reaction1, params.data(), custom_ctx);
    double val[128];
    for (int i = 0; i < S; ++i) val[i] = (double)state[i];
    for (int i = 0; i < P; ++i) val[S + i] = params[i];
    return evaluate_program(programs[reaction1 - 1], val);
}

namespace {
inline bool same_bits(double a, double b) { return std::memcmp(&a, &b, sizeof a) == 0; }
// splitmix64: the probe must not depend on the C library's generator (the SSA streams are Philox; this one only picks probe states)
inline uint64_t probe_next(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
}  // namespace

bool probe_custom(const HostModel& m, int32_t max_molecules, CustomProbe& out, int nverify) {
    out = CustomProbe();
    out.species.assign((size_t)m.R, -2);
    out.sa.assign((size_t)m.R, -1);
    out.sb.assign((size_t)m.R, -1);
    out.coef.assign((size_t)m.R, 0.0);
    if (!m.custom || m.S < 1 || m.R < 1 || max_molecules < 1) return false;
    const int S = m.S, R = m.R;
    const int64_t tlen = (int64_t)max_molecules + 1;
    uint64_t rng = 0x6B66737042323030ull;
    auto call = [&](const int32_t* st, int k) { return m.custom(st, k + 1, m.params.data(), m.custom_ctx); };
    // base states: small counts (where Hill terms and combinatorial factors are most curved), all ones, and a few spread ones
    std::vector<std::vector<int32_t>> bases;
    bases.emplace_back((size_t)S, 1);
    bases.emplace_back((size_t)S, 2);
    for (int b = 0; b < 6; ++b) {
        std::vector<int32_t> st((size_t)S);
        const int32_t range = b < 3 ? std::min<int32_t>(max_molecules, 12) : std::min<int32_t>(max_molecules, 300);
        for (int s = 0; s < S; ++s) st[s] = 1 + (int32_t)(probe_next(rng) % (uint64_t)range);
        bases.push_back(st);
    }
    std::vector<int32_t> line;
    for (int32_t c = 0; c <= std::min<int32_t>(max_molecules, 40); ++c) line.push_back(c);
    for (int q = 0; q < 8; ++q) line.push_back((int32_t)(probe_next(rng) % (uint64_t)tlen));
    line.push_back(max_molecules);
    std::vector<int32_t> st((size_t)S);
    std::vector<uint32_t> masks((size_t)R, 0);
    bool supported = true;
    for (int k = 0; k < R; ++k) {
        uint32_t mask = 0;
        for (const auto& base : bases) {
            const double a0 = call(base.data(), k);
            for (int s = 0; s < S; ++s) {
                if (mask & (1u << s)) continue;
                st = base;
                for (int32_t c : line) {
                    st[s] = c;
                    if (!same_bits(call(st.data(), k), a0)) { mask |= 1u << s; break; }
                }
            }
        }
        masks[k] = mask;
        const int nsp = __builtin_popcount(mask);
        if (nsp <= 1) {
            int sp = 0;
            while (mask > 1) { mask >>= 1; ++sp; }
            out.species[k] = sp;
        } else if (nsp == 2) {
            out.species[k] = -2;                                   // candidate for the bilinear form, decided below
        } else {
            out.species[k] = -2;
            supported = false;
        }
    }
    out.all_single = true;
    for (int k = 0; k < R; ++k) out.all_single = out.all_single && out.species[k] >= 0;
    if (!supported) return false;
    // tables of the single-species reactions; coefficient and operand order of the two-species ones:
    // a(x) = fl(fl(c * x_a) * x_b) with c = a(x_a = 1, x_b = 1) -- mass action as a compiler evaluates c*x*y, either operand first
    out.tables.assign((size_t)R * tlen, 0.0);
    for (int k = 0; k < R; ++k) {
        std::fill(st.begin(), st.end(), 0);
        if (out.species[k] >= 0) {
            for (int64_t c = 0; c < tlen; ++c) {
                st[out.species[k]] = (int32_t)c;
                out.tables[(size_t)k * tlen + c] = call(st.data(), k);
            }
            continue;
        }
        int p = -1, q = -1;
        for (int s = 0; s < S; ++s)
            if (masks[k] & (1u << s)) { if (p < 0) p = s; else q = s; }
        st[p] = 1; st[q] = 1;
        const double c = call(st.data(), k);
        out.coef[k] = c;
        // which operand is multiplied first: decide on a grid of small counts, then verify on the random states below
        bool pq = true, qp = true;
        for (int32_t u = 0; u <= 24 && (pq || qp); ++u)
            for (int32_t v = 0; v <= 24; ++v) {
                st[p] = u * 37 % (max_molecules + 1); st[q] = v * 91 % (max_molecules + 1);
                const double a = call(st.data(), k);
                pq = pq && same_bits(a, (c * (double)st[p]) * (double)st[q]);
                qp = qp && same_bits(a, (c * (double)st[q]) * (double)st[p]);
            }
        if (pq) { out.sa[k] = p; out.sb[k] = q; }
        else if (qp) { out.sa[k] = q; out.sb[k] = p; }
        else return false;                                         // reads two species in some other way: host path
    }
    // verification: the callback against the tables / the bilinear form on states drawn at three scales, bit for bit
    for (int t = 0; t < nverify; ++t) {
        const int32_t range = t % 3 == 0 ? std::min<int32_t>(max_molecules, 16) : t % 3 == 1 ? std::min<int32_t>(max_molecules, 400) : max_molecules;
        for (int s = 0; s < S; ++s) st[s] = (int32_t)(probe_next(rng) % (uint64_t)(range + 1));
        for (int k = 0; k < R; ++k) {
            const double a = call(st.data(), k);
            const double e = out.species[k] >= 0 ? out.tables[(size_t)k * tlen + st[out.species[k]]]
                                                 : (out.coef[k] * (double)st[out.sa[k]]) * (double)st[out.sb[k]];
            if (!same_bits(a, e)) {
                if (out.species[k] >= 0) { out.species[k] = -2; out.all_single = false; }
                out.sa[k] = out.sb[k] = -1;
                return false;
            }
        }
    }
    out.ok = true;
    return true;
}

bool probe_custom_single_species(const HostModel& m, int32_t max_molecules, std::vector<int32_t>& species, std::vector<double>& tables,
                                 int nverify) {
    CustomProbe pr;
    const bool ok = probe_custom(m, max_molecules, pr, nverify);
    species = pr.species;
    tables.clear();
    if (ok && pr.all_single) { tables = pr.tables; return true; }
    return false;
}

bool parse_reaction(const std::string& line, const std::vector<std::string>& species, int32_t* vec, std::string& err) {
    std::istringstream is(line);
    std::vector<std::string> terms;
    std::string w;
    int direction = 0;
    size_t nleft = 0;
    while (is >> w) {
        if (w == "->") { direction = 1; nleft = terms.size(); }
        else if (w == "<-") { direction = 2; nleft = terms.size(); }
        else if (w != "+") terms.push_back(w);
    }
    if (direction == 0) { err = "SYNTAX ERROR IN CHEMICAL REACTION, ONLY ONE SIDE WAS WRITTEN: " + line; return false; }
    const size_t S = species.size();
    for (size_t j = 0; j < S; ++j) vec[j] = 0;
    int coeff = 0;                    // deliberately carried over between species/terms
    for (size_t i = 0; i < terms.size(); ++i) {
        const std::string& t = terms[i];
        if (t == "0") continue;
        for (size_t j = 0; j < S; ++j) {
            size_t k = t.find(species[j]);
            if (k == std::string::npos) coeff = 0;
            else if (t.compare(k, std::string::npos, species[j]) == 0) coeff = k > 0 ? std::atoi(t.substr(0, k).c_str()) : 1;
            if (i < nleft) vec[j] -= coeff; else vec[j] += coeff;
        }
    }
    if (direction == 2) for (size_t j = 0; j < S; ++j) vec[j] = -vec[j];
    return true;
}

bool load_model_file(const std::string& path, HostModel& m, std::string& err) {
    std::ifstream in(path);
    if (!in) { err = "ERROR OPENING FILE " + path; return false; }
    std::vector<std::string> lines;
    for (std::string ln; std::getline(in, ln);) {
        if (!ln.empty() && ln.back() == '\r') ln.pop_back();
        lines.push_back(ln);
    }
    size_t pos = 0;
    auto next_nonblank = [&](std::string& out) {
        while (pos < lines.size()) {
            const std::string& ln = lines[pos++];
            if (!blank(ln)) { out = ln; return true; }
        }
        return false;
    };
    bool have_species = false, have_params = false, have_ns = false, have_nr = false, have_np = false;
    std::string ln;
    while (next_nonblank(ln)) {
        std::string key = upper(first_token(ln));
        std::string v;
        if (key == "NSPECIES") { if (!next_nonblank(v)) break; m.S = std::atoi(first_token(v).c_str()); have_ns = true; }
        else if (key == "NREACTIONS") { if (!next_nonblank(v)) break; m.R = std::atoi(first_token(v).c_str()); have_nr = true; }
        else if (key == "NPARAMETERS") { if (!next_nonblank(v)) break; m.P = std::atoi(first_token(v).c_str()); have_np = true; }
        else if (key == "SPECIES") {
            if (!have_ns) { err = "MODEL INPUT ERROR: NUMBER OF SPECIES NOT DECLARED."; return false; }
            m.species.clear();
            for (int i = 0; i < m.S; ++i) { if (!next_nonblank(v)) { err = "MODEL INPUT ERROR: TOO FEW SPECIES NAMES."; return false; } m.species.push_back(first_token(v)); }
            have_species = true;
        } else if (key == "PARAMETERS") {
            if (!have_np) { err = "MODEL INPUT ERROR: NUMBER OF PARAMETERS NOT DECLARED BEFORE SPECIFYING PARAMETER NAMES."; return false; }
            m.parameters.clear();
            for (int i = 0; i < m.P; ++i) { if (!next_nonblank(v)) { err = "MODEL INPUT ERROR: TOO FEW PARAMETER NAMES."; return false; } m.parameters.push_back(first_token(v)); }
            m.params.assign(m.P, 0.0);
            have_params = true;
        } else if (key == "REACTIONS") {
            if (!have_species) { err = "MODEL INPUT ERROR: REACTIONS STATED BEFORE SPECIES NAMES ARE DECLARED."; return false; }
            if (!have_nr) { err = "MODEL INPUT ERROR: NUMBER OF REACTIONS NOT DECLARED."; return false; }
            m.stoich.assign((size_t)m.S * m.R, 0);
            for (int k = 0; k < m.R; ++k) {
                if (pos >= lines.size()) { err = "MODEL INPUT ERROR: FEWER REACTION LINES THAN NREACTIONS."; return false; }
                if (!parse_reaction(lines[pos++], m.species, &m.stoich[(size_t)k * m.S], err)) return false;
            }
        } else if (key == "PROPENSITIES") {
            if (!have_species || !have_params) { err = "MODEL INPUT ERROR: PROPENSITIES SPECIFIED BEFORE ALL SPECIES AND PARAMETERS ARE NAMED."; return false; }
            std::vector<std::string> vars = m.species;
            vars.insert(vars.end(), m.parameters.begin(), m.parameters.end());
            m.programs.assign(m.R, Program());
            m.propensity_strings.assign(m.R, "");
            for (int k = 0; k < m.R; ++k) {
                if (pos >= lines.size()) { err = "MODEL INPUT ERROR: FEWER PROPENSITY LINES THAN NREACTIONS."; return false; }
                m.propensity_strings[k] = lines[pos++];
                if (!compile_expression(m.propensity_strings[k], vars, m.programs[k], err)) return false;
            }
        }
    }
    if (!have_ns || !have_nr) { err = "MODEL INPUT ERROR: NSPECIES/NREACTIONS MISSING."; return false; }
    if (m.programs.empty()) m.programs.assign(m.R, Program());
    if (m.params.empty()) m.params.assign(m.P > 0 ? m.P : 0, 0.0);
    if (m.stoich.empty()) m.stoich.assign((size_t)m.S * m.R, 0);
    return true;
}

}  // namespace kfsp
