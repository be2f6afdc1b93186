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
bool load_model_file(const std::string& path, HostModel& m, std::string& err);
bool parse_reaction(const std::string& line, const std::vector<std::string>& species, int32_t* vec, std::string& err);

}  // namespace kfsp
