"""ORACLE (test infrastructure only -- never imported by the product path).

Pure-Python restatement of the reference's `.input` model reader:

  CME_MODEL%LOAD   /root/reference/src/model/ModelModule.f90:59-161
  STOICH_INPUT     /root/reference/src/model/ModelModule.f90:219-297

Documented deviation: the reference compares section keywords
against upper-case literals while the shipped files are lower case, so as shipped it
cannot read its own models.  Keywords are matched case-insensitively here.

STOICH_INPUT quirks that are kept: species are matched by *substring* (INDEX) and a
partial match leaves COEFF at its previous value (ModelModule.f90:274-284) -- for the
shipped models the stale value is always 0, which the assertion below checks.
"""
from . import fparser


class InputModel:
    def __init__(self):
        self.nspecies = 0
        self.nreactions = 0
        self.nparameters = 0
        self.species_names = []
        self.parameter_names = []
        self.stoichiometry = []      # [reaction][species]
        self.propensity_strings = []
        self.programs = []           # fparser.Program per reaction


def stoich_input(nspecies, line, species_names):
    """ModelModule.f90:219-297; returns the stoichiometry vector of one reaction."""
    terms = []
    nleft = 0
    direction = 0
    for word in line.split():        # blank-separated words (:246-262)
        if word == "->":
            direction = 1
            nleft = len(terms)
        elif word == "<-":
            direction = 2
            nleft = len(terms)
        elif word != "+":
            terms.append(word)
    if direction == 0:
        raise ValueError("SYNTAX ERROR IN CHEMICAL REACTION, ONLY ONE SIDE WAS WRITTEN.")
    vec = [0] * nspecies
    coeff = 0                        # stale across iterations, like the Fortran local
    for i, term in enumerate(terms):
        if term == "0":
            continue
        for j in range(nspecies):
            name = species_names[j]
            k = term.find(name)      # INDEX(): first occurrence, -1 if absent
            if k < 0:
                coeff = 0
            elif term[k:] == name:
                coeff = int(term[:k]) if k > 0 else 1
            # else: partial match, COEFF keeps its previous value (:278-284)
            if i < nleft:
                vec[j] -= coeff
            else:
                vec[j] += coeff
    if direction == 2:
        vec = [-v for v in vec]
    return vec


def load(path):
    """ModelModule.f90:59-161."""
    m = InputModel()
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    pos = 0

    def next_nonblank():
        nonlocal pos
        while pos < len(lines):
            ln = lines[pos]
            pos += 1
            if ln.strip():
                return ln
        return None

    while True:
        ln = next_nonblank()
        if ln is None:
            break
        key = ln.split()[0].upper()
        if key == "NSPECIES":
            m.nspecies = int(next_nonblank().split()[0])
        elif key == "NREACTIONS":
            m.nreactions = int(next_nonblank().split()[0])
        elif key == "NPARAMETERS":
            m.nparameters = int(next_nonblank().split()[0])
        elif key == "SPECIES":
            m.species_names = [next_nonblank().split()[0] for _ in range(m.nspecies)]
        elif key == "PARAMETERS":
            m.parameter_names = [next_nonblank().split()[0] for _ in range(m.nparameters)]
        elif key == "REACTIONS":
            if not m.species_names:
                raise ValueError("MODEL INPUT ERROR: REACTIONS STATED BEFORE SPECIES NAMES ARE DECLARED.")
            for _ in range(m.nreactions):
                if pos >= len(lines):
                    raise ValueError("MODEL INPUT ERROR: FEWER REACTION LINES THAN NREACTIONS.")
                ln = lines[pos]      # READ(10,'(A)') takes the very next record
                pos += 1
                m.stoichiometry.append(stoich_input(m.nspecies, ln, m.species_names))
        elif key == "PROPENSITIES":
            if not m.species_names or not m.parameter_names:
                raise ValueError("MODEL INPUT ERROR: PROPENSITIES SPECIFIED BEFORE ALL SPECIES AND PARAMETERS ARE NAMED.")
            fvar = list(m.species_names) + list(m.parameter_names)
            for _ in range(m.nreactions):
                if pos >= len(lines):
                    raise ValueError("MODEL INPUT ERROR: FEWER PROPENSITY LINES THAN NREACTIONS.")
                ln = lines[pos]
                pos += 1
                m.propensity_strings.append(ln.strip())
                m.programs.append(fparser.compile_expression(ln.strip(), fvar))
    return m
