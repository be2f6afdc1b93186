// Host-side mirror of the reference's CME_MODEL (src/model/ModelModule.f90:14-42):
// sizes, stoichiometry, parameter values and one compiled propensity program per
// reaction.  The byte code is the reference parser's (src/parser/FortranParser.f90:52-73)
// so that a Fortran host can hand over its PROPPARSER arrays unchanged.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/kfsp.h"

namespace kfsp {

enum Op : int32_t {
    cImmed = 1, cNeg, cAdd, cSub, cMul, cDiv, cPow, cAbs, cExp, cLog10, cLog, cSqrt, cSinh, cCosh, cTanh,
    cSin, cCos, cTan, cAsin, cAcos, cAtan, VarBegin
};

struct Program {
    std::vector<int32_t> code;
    std::vector<double> immed;
    int stack_depth = 0;
    bool empty() const { return code.empty(); }
};

struct HostModel {
    int32_t S = 0, R = 0, P = 0;
    std::vector<int32_t> stoich;            // S*R, species fastest
    std::vector<double> params;             // P
    std::vector<std::string> species, parameters, propensity_strings;
    std::vector<Program> programs;          // R
    kfsp_propensity_fn custom = nullptr;
    void* custom_ctx = nullptr;
    std::string error;

    double propensity(const int32_t* state, int reaction1) const;
};

// Compile `expr` over the variable names `vars`; returns false and sets err on a syntax error.
bool compile_expression(const std::string& expr, const std::vector<std::string>& vars, Program& out, std::string& err);
// Depth of the evaluation stack a program needs, -1 if malformed.
int program_stack_depth(const Program& p, int nvars);
double evaluate_program(const Program& p, const double* val);
// Which species a program reads (bit mask over the first S variables) and whether it contains an
// operation that is not correctly rounded on every platform (pow, exp, log, trigonometric ...).
void program_profile(const Program& p, int S, uint32_t* species_mask, bool* inexact);
// evaluate_program with the interpreter's early exits reported: *aborted is set when a division by zero or a domain
// error (log of a non-positive number, sqrt of a negative one, asin/acos out of range) ended the evaluation with 0.
double evaluate_program_checked(const Program& p, const double* val, bool* aborted);

// Partial evaluation of a propensity program for the index-only SpMV (spmv_variant = 2): every maximal sub-expression
// that reads AT MOST ONE species becomes a `term` (tabulated on the host over that species' count, same libm and same
// operation order as the whole program), and what is left above the terms -- the nodes that combine several species --
// must consist of + - * and unary minus only (IEEE-exact operations the device reproduces bit for bit).  c5*DNA*D
// becomes T1[DNA] * T2[D] with T1 = c5*DNA, T2 = D.  ops is the postfix program over the terms: a value t >= 0 pushes
// term t, a negative value -op applies the byte-code operation op (cNeg, cAdd, cSub, cMul).
struct FactoredTerm {
    int species = -1;                        // -1: reads no species (a constant)
    Program prog;                            // the sub-expression, variables numbered as in the full program
};
struct Factored {
    std::vector<FactoredTerm> terms;
    std::vector<int32_t> ops;
};
// false if the program is malformed or a multi-species node is not one of + - * neg
bool factor_program(const Program& p, int S, Factored& out);

// CUSTOMPROP structure by probing (the callback is opaque): for every reaction, vary one species at a time around several
// base states and record which ones change the value.  If every reaction reads AT MOST ONE species, tabulate it over the
// counts 0..max_molecules (tables[k*(max_molecules+1) + c], species[k] = that species, 0 for a constant) and verify the
// tables bit for bit against the callback on `nverify` pseudo-random states; true only if all of that holds.  A model that
// passes can be served from device tables like a parsed single-species propensity (no host round trips, every SpMV variant,
// every multi-GPU layout); one that does not stays on the host-callback path.  species[k] = -2 marks a reaction that reads
// several species (reported even when the function returns false).
// The general probe: besides single-species reactions it recognises BILINEAR mass action, a(x) = fl(fl(c * x_a) * x_b) with
// c = a(x_a = 1, x_b = 1) (examples/transcr6d.f90:74,78: parameters(5) * state(DNA) * state(D)), in either operand order, again
// verified bit for bit on the random states.  ok: every reaction is of one of the two kinds, i.e. the model can be evaluated on
// the device (tables for the single-species reactions, the three-operation program c * X_a * X_b for the bilinear ones).
struct CustomProbe {
    std::vector<int32_t> species;      // per reaction: the species of a single-species reaction (0 for a constant), -2 otherwise
    std::vector<int32_t> sa, sb;       // bilinear reactions: first and second operand species, else -1
    std::vector<double> coef;          // bilinear reactions: c
    std::vector<double> tables;        // [k*(max_molecules+1) + count] for the single-species reactions
    bool all_single = false;
    bool ok = false;
};
bool probe_custom(const HostModel& m, int32_t max_molecules, CustomProbe& out, int nverify = 16384);
bool probe_custom_single_species(const HostModel& m, int32_t max_molecules, std::vector<int32_t>& species, std::vector<double>& tables,
                                 int nverify = 16384);

bool load_model_file(const std::string& path, HostModel& m, std::string& err);
bool parse_reaction(const std::string& line, const std::vector<std::string>& species, int32_t* vec, std::string& err);

}  // namespace kfsp
