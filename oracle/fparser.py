"""ORACLE (test infrastructure only -- never imported by the product path).

Pure-Python restatement of the reference's expression compiler, Roland Schmehl's
fparser as shipped in /root/reference/src/parser/FortranParser.f90.  Only the pieces
that decide *evaluation order* (and therefore last-bit values of a propensity) are
restated:

  parse            FortranParser.f90:172-184   ('**' -> '^ ', strip blanks)
  CompileSubstr    FortranParser.f90:627-723   (right-to-left operator split, + - * / ^)
  IsBinaryOp       FortranParser.f90:726-765
  MathItemIndex    FortranParser.f90:580-598
  RealNum          FortranParser.f90:768-841
  evaluate         FortranParser.f90:187-302   (stack machine, 21 opcodes)

The byte code numbering is the reference's (FortranParser.f90:52-73): cImmed=1 ...
cAtan=21, variables start at VarBegin=22.

Pinned by tests/test_oracle_parser.py against the closed-form PROP function of
/root/reference/test/TestModelParser.f90:80-102 (the only result-bearing check the
reference has for this code).
"""
import math

cImmed, cNeg, cAdd, cSub, cMul, cDiv, cPow = 1, 2, 3, 4, 5, 6, 7
cAbs, cExp, cLog10, cLog, cSqrt, cSinh, cCosh, cTanh = 8, 9, 10, 11, 12, 13, 14, 15
cSin, cCos, cTan, cAsin, cAcos, cAtan = 16, 17, 18, 19, 20, 21
VarBegin = 22

OPS = {cAdd: "+", cSub: "-", cMul: "*", cDiv: "/", cPow: "^"}
FUNCS = {cAbs: "abs", cExp: "exp", cLog10: "log10", cLog: "log", cSqrt: "sqrt",
         cSinh: "sinh", cCosh: "cosh", cTanh: "tanh", cSin: "sin", cCos: "cos",
         cTan: "tan", cAsin: "asin", cAcos: "acos", cAtan: "atan"}


class ParseError(ValueError):
    pass


def _math_function_index(s):
    """FortranParser.f90:440-459: first function (in opcode order) whose name is a
    case-insensitive prefix of s."""
    for code in range(cAbs, cAtan + 1):
        name = FUNCS[code]
        k = min(len(name), len(s))
        # the Fortran compares str(1:k) blank-padded to LEN(Funcs)=5 with Funcs(j)
        if s[:k].lower().ljust(5) == name.ljust(5):
            return code
    return 0


def _real_num(s):
    """FortranParser.f90:768-841.  Returns (value, inext0) where inext0 is the 0-based
    index of the first character after the number, or raises ParseError."""
    Bflag, InMan, Pflag, Eflag, InExp = True, False, False, False, False
    DInMan = DInExp = False
    ib = 0
    i = 0
    n = len(s.rstrip(" "))
    while i < n:
        c = s[i]
        if c == " ":
            ib += 1
            if InMan or Eflag or InExp:
                break
        elif c in "+-":
            if Bflag:
                InMan, Bflag = True, False
            elif Eflag:
                InExp, Eflag = True, False
            else:
                break
        elif c.isdigit():
            if Bflag:
                InMan, Bflag = True, False
            elif Eflag:
                InExp, Eflag = True, False
            if InMan:
                DInMan = True
            if InExp:
                DInExp = True
        elif c == ".":
            if Bflag:
                Pflag = True
                InMan, Bflag = True, False
            elif InMan and not Pflag:
                Pflag = True
            else:
                break
        elif c in "eEdD":
            if InMan:
                Eflag, InMan = True, False
            else:
                break
        else:
            break
        i += 1
    err = (ib > i - 1) or (not DInMan) or ((Eflag or InExp) and not DInExp)
    if err:
        raise ParseError("invalid number format: %r" % s)
    txt = s[ib:i].replace("d", "e").replace("D", "e")
    return float(txt), i


def _is_binary_op(j, F):
    """FortranParser.f90:726-765 (j is a 0-based position in F)."""
    res = True
    if F[j] in "+-":
        if j == 0:
            res = False
        elif F[j - 1] in "+-*/^(":
            res = False
        elif j + 1 < len(F) and F[j + 1].isdigit() and F[j - 1] in "eEdD":
            Dflag = Pflag = False
            k = j - 1
            while k > 0:
                k -= 1
                if F[k].isdigit():
                    Dflag = True
                elif F[k] == ".":
                    if Pflag:
                        break
                    Pflag = True
                else:
                    break
            if Dflag and (k == 0 or F[k] in "+-*/^("):
                res = False
    return res


def _completely_enclosed(F, b, e):
    """FortranParser.f90:601-624 (b, e inclusive 0-based)."""
    if b > e or F[b] != "(" or F[e] != ")":
        return False
    k = 0
    for j in range(b + 1, e):
        if F[j] == "(":
            k += 1
        elif F[j] == ")":
            k -= 1
        if k < 0:
            break
    return k == 0


class Program:
    """Compiled propensity: byte code + immediates, as EquationParser holds them."""

    def __init__(self, expr, variables):
        self.orig = expr
        self.variables = [v.strip() for v in variables]
        # parse(): Replace('**','^ ') then RemoveSpaces
        self.F = expr.replace("**", "^ ").replace(" ", "").replace("\t", "")
        if not self.F:
            raise ParseError("empty expression")
        self.code = []
        self.immed = []
        self.stack_ptr = 0
        self.stack_size = 0
        self._compile(0, len(self.F) - 1)

    # -- FortranParser.f90:462-492
    def _variable_index(self, s):
        i = 0
        while i < len(s) and s[i] not in "+-*/^) ":
            i += 1
        name = s[:i]
        for j, v in enumerate(self.variables):
            if name == v:
                return j + 1
        return 0

    # -- FortranParser.f90:580-598
    def _math_item(self, b, e):
        F = self.F
        if F[b] in "0123456789.":
            val, _ = _real_num(F[b:e + 1])
            self.immed.append(val)
            return cImmed
        n = self._variable_index(F[b:e + 1])
        if n == 0:
            raise ParseError("invalid element %r in %r" % (F[b:e + 1], self.orig))
        return VarBegin + n - 1

    # -- FortranParser.f90:627-723
    def _compile(self, b, e):
        F = self.F
        if b > e:
            raise ParseError("missing operand in %r" % self.orig)
        if F[b] == "+":
            return self._compile(b + 1, e)
        if _completely_enclosed(F, b, e):
            return self._compile(b + 1, e - 1)
        if F[b].isalpha() and F[b].isascii():
            n = _math_function_index(F[b:e + 1])
            if n > 0:
                p = F.find("(", b, e + 1)
                if p >= 0 and _completely_enclosed(F, p, e):
                    self._compile(p + 1, e - 1)
                    self.code.append(n)
                    return
        elif F[b] == "-":
            if _completely_enclosed(F, b + 1, e):
                self._compile(b + 2, e - 1)
                self.code.append(cNeg)
                return
            if b + 1 <= e and F[b + 1].isalpha() and F[b + 1].isascii():
                n = _math_function_index(F[b + 1:e + 1])
                if n > 0:
                    p = F.find("(", b + 1, e + 1)
                    if p >= 0 and _completely_enclosed(F, p, e):
                        self._compile(p + 1, e - 1)
                        self.code.append(n)
                        self.code.append(cNeg)
                        return
        for io in range(cAdd, cPow + 1):
            k = 0
            for j in range(e, b - 1, -1):
                if F[j] == ")":
                    k += 1
                elif F[j] == "(":
                    k -= 1
                if k == 0 and F[j] == OPS[io] and _is_binary_op(j, F):
                    if F[j] in "*/^" and F[b] == "-":
                        self._compile(b + 1, e)
                        self.code.append(cNeg)
                        return
                    self._compile(b, j - 1)
                    self._compile(j + 1, e)
                    self.code.append(io)
                    self.stack_ptr -= 1
                    return
        b2 = b + 1 if F[b] == "-" else b
        self.code.append(self._math_item(b2, e))
        self.stack_ptr += 1
        if self.stack_ptr > self.stack_size:
            self.stack_size += 1
        if b2 > b:
            self.code.append(cNeg)

    # -- FortranParser.f90:187-302
    def evaluate(self, val):
        st = []
        dp = 0
        for op in self.code:
            if op == cImmed:
                st.append(self.immed[dp]); dp += 1
            elif op == cNeg:
                st[-1] = -st[-1]
            elif op == cAdd:
                y = st.pop(); st[-1] = st[-1] + y
            elif op == cSub:
                y = st.pop(); st[-1] = st[-1] - y
            elif op == cMul:
                y = st.pop(); st[-1] = st[-1] * y
            elif op == cDiv:
                if st[-1] == 0.0:
                    return 0.0
                y = st.pop(); st[-1] = st[-1] / y
            elif op == cPow:
                y = st.pop(); st[-1] = _fpow(st[-1], y)
            elif op == cAbs:
                st[-1] = abs(st[-1])
            elif op == cExp:
                st[-1] = math.exp(st[-1])
            elif op == cLog10:
                if st[-1] <= 0.0:
                    return 0.0
                st[-1] = math.log10(st[-1])
            elif op == cLog:
                if st[-1] <= 0.0:
                    return 0.0
                st[-1] = math.log(st[-1])
            elif op == cSqrt:
                if st[-1] < 0.0:
                    return 0.0
                st[-1] = math.sqrt(st[-1])
            elif op == cSinh:
                st[-1] = math.sinh(st[-1])
            elif op == cCosh:
                st[-1] = math.cosh(st[-1])
            elif op == cTanh:
                st[-1] = math.tanh(st[-1])
            elif op == cSin:
                st[-1] = math.sin(st[-1])
            elif op == cCos:
                st[-1] = math.cos(st[-1])
            elif op == cTan:
                st[-1] = math.tan(st[-1])
            elif op == cAsin:
                if st[-1] < -1.0 or st[-1] > 1.0:
                    return 0.0
                st[-1] = math.asin(st[-1])
            elif op == cAcos:
                if st[-1] < -1.0 or st[-1] > 1.0:
                    return 0.0
                st[-1] = math.acos(st[-1])
            elif op == cAtan:
                st[-1] = math.atan(st[-1])
            else:
                st.append(val[op - VarBegin])
        return st[0]


def _fpow(a, b):
    """Fortran a**b for two reals == C pow(a, b) (no Python exceptions)."""
    try:
        return math.pow(a, b)
    except (OverflowError, ValueError):
        if a == 0.0 and b < 0:
            return math.inf
        if a < 0:
            return math.nan
        return math.inf


def compile_expression(expr, variables):
    p = Program(expr, variables)
    return p
