"""Pin the CPU oracle (and, through tests/, the CUDA path) to the reference's own SOURCE TEXT.

The reference is Fortran 90 and cannot be compiled in this image (no f951), and it stores no golden vectors
for the EVP path.  Its hot-path subroutines, however, are straight-line fp64 arithmetic:

  B grid   stress            cicecore/cicedyn/dynamics/ice_dyn_evp.F90      
           stepu             cicecore/cicedyn/dynamics/ice_dyn_shared.F90   (a3)
           strain_rates      .../ice_dyn_shared.F90                         (a4)
           visc_replpress    .../ice_dyn_shared.F90                         (a5)
  C grid   strain_rates_U, strain_rates_Tdt, stepu_C, stepv_C   ice_dyn_shared.F90          (a12)
           stressC_T, stressC_U, div_stress_Ex, div_stress_Ny   ice_dyn_evp.F90
           grid_average_X2Y_1 / X2YS / X2YA                     cicecore/cicedyn/infrastructure/ice_grid.F90
  CD grid  stressCD_T, stressCD_U, div_stress_Ey, div_stress_Nx  ice_dyn_evp.F90           (a13)
           strain_rates_Tdtsd, stepuv_CD                         ice_dyn_shared.F90
  next     deformations      .../ice_dyn_shared.F90                         (8f rank 2)
  constants                  cicecore/shared/ice_constants.F90              (p111 = c1/c9 ...)

This script READS those subroutines from /root/reference at generation time, transliterates them statement
by statement into Python -- same operators, same operand order, same parentheses; a Python float is an IEEE
double and `a op b` is rounded once, exactly like Fortran without reassociation or FMA contraction -- executes
the result on synthetic cases inside a driver that repeats the reference's loop (ice_dyn_evp.F90:859-913 and
:936-1101: call order, index lists, halo points), and writes the outputs as golden vectors.  Nothing of the
arithmetic is restated by hand here: the translator knows Fortran syntax (continuations, do / if / select case,
call with intent(out) scalars, generic interfaces, array references and sections, derived-type members, `x**2`),
not the formulas.  The halo between the calls is not transliterated (ice_boundary.F90 is MPI code): the B-grid
driver applies the single-block cyclic wrap directly, the C-grid driver is handed the oracle's halo routine.

Consumers: tests/test_oracle.py, tests/test_cgrid.py (the C oracle against the committed vectors, bit for bit,
anywhere; re-derivation from /root/reference where it exists) and tests/test_gpu_parity.py, tests/test_cgrid.py
-m gpu (the CUDA path against the same vectors).

usage:  python tests/golden/ref_translit.py [--write] [--show]     (repo root, in the container that has /root/reference)
"""
import math
import os
import re
import sys

import numpy as np

REF = os.environ.get("CICE_REFERENCE", "/root/reference")
F_EVP = "cicecore/cicedyn/dynamics/ice_dyn_evp.F90"
F_SHARED = "cicecore/cicedyn/dynamics/ice_dyn_shared.F90"
F_CONST = "cicecore/shared/ice_constants.F90"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_source_vectors.npz")


# ------------------------------------------------------------------------------------------
# Fortran source -> logical lines
# ------------------------------------------------------------------------------------------
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def logical_lines(text):
    """join `&` continuations, drop comments and blank lines, keep original case."""
    res, cur = [], ""
    for raw in text.splitlines():
        if raw.lstrip().startswith("#"):  # cpp
            continue
        s = _strip_comment(raw).strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        res.append(cur + s)
        cur = ""
    return res


def subroutine_lines(path, name):
    lines = logical_lines(open(os.path.join(REF, path)).read())
    start = None
    for n, ln in enumerate(lines):
        if re.match(rf"^\s*subroutine\s+{name}\b", ln, re.I):
            start = n
        if start is not None and re.match(rf"^\s*end\s+subroutine\s+{name}\b", ln, re.I):
            return lines[start:n + 1]
    raise RuntimeError(f"subroutine {name} not found in {path}")


# ------------------------------------------------------------------------------------------
# expression translator (precedence of Fortran 90: ** > * / > unary +- > binary +- > relational > .not. > .and. > .or.)
# ------------------------------------------------------------------------------------------
TOK = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?(?:_\w+)?)
  | (?P<dot>\.(?:and|or|not|true|false|eq|ne|lt|le|gt|ge)\.)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<str>'[^']*'|"[^"]*")
  | (?P<op>\*\*|==|/=|<=|>=|[-+*/(),<>=:%])
)""", re.X | re.I)

INTRINSICS = {"sqrt": "math.sqrt", "max": "max", "min": "min", "abs": "abs", "sign": "_sign", "trim": "_trim", "real": "float",
              "present": "_present"}
REL = {"==": "==", "/=": "!=", "<": "<", ">": ">", "<=": "<=", ">=": ">=",
       ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


def tokenize(s):
    pos, toks = 0, []
    s = s.rstrip()
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m or m.end() == pos:
            raise SyntaxError(f"cannot tokenize {s[pos:]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    return toks


class Expr:
    def __init__(self, toks, arrays):
        self.t, self.p, self.arrays = toks, 0, arrays

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.p += 1
        return tok

    def expect(self, v):
        k, x = self.next()
        if x != v:
            raise SyntaxError(f"expected {v!r}, got {x!r}")

    def parse(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] and self.peek()[1].lower() == ".or.":
            self.next()
            a = f"({a} or {self.p_and()})"
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] and self.peek()[1].lower() == ".and.":
            self.next()
            a = f"({a} and {self.p_not()})"
        return a

    def p_not(self):
        if self.peek()[1] and self.peek()[1].lower() == ".not.":
            self.next()
            return f"(not {self.p_not()})"
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        v = self.peek()[1]
        if v and v.lower() in REL:
            self.next()
            return f"({a} {REL[v.lower()]} {self.p_add()})"
        return a

    def p_add(self):
        v = self.peek()[1]
        if v in ("+", "-"):           # leading sign applies to the first TERM: -a*b == -(a*b)
            self.next()
            a = f"({v}{self.p_mul()})"
        else:
            a = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = f"({a} {op} {self.p_mul()})"
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            a = f"({a} {op} {self.p_pow()})"
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek()[1] == "**":
            self.next()
            b = self.p_pow() if self.peek()[1] != "-" else None
            if b is None:
                raise SyntaxError("negative exponent not supported")
            if b in ("2", "2.0"):
                return f"_sq({a})"     # x**2 is x*x in every Fortran compiler; Python's pow() is not guaranteed to be
            if b in ("4", "4.0"):
                return f"_sq(_sq({a}))"  # x**4 = (x*x)*(x*x) by repeated squaring
            raise SyntaxError(f"exponent {b} not supported")
        return a

    def p_primary(self):
        k, v = self.next()
        if k == "num":
            v = re.sub(r"_\w+$", "", v)
            v = re.sub(r"[dD]", "e", v)
            return v if ("." in v or "e" in v.lower()) else v  # integers stay integers (indices, exponents)
        if k == "dot":
            return {".true.": "True", ".false.": "False"}[v.lower()]
        if k == "str":
            return repr(v[1:-1])
        if k == "op" and v == "(":
            e = self.parse()
            self.expect(")")
            return f"({e})"
        if k == "id":
            v = v.lower()                         # Fortran is case-insensitive (deltaU / DeltaU)
            while self.peek()[1] == "%":          # derived-type member: this_block%ilo
                self.next()
                v = f"{v}.{self.next()[1].lower()}"
            if self.peek()[1] == "(":
                # array section name(:,:) / name(:) -> the whole array
                q = self.p + 1
                sec = True
                while q < len(self.t) and self.t[q][1] != ")":
                    if self.t[q][1] not in (":", ","):
                        sec = False
                        break
                    q += 1
                if sec and q < len(self.t) and q > self.p + 1:
                    self.p = q + 1
                    return v
                self.next()
                args = []
                if self.peek()[1] != ")":
                    args.append(self.parse())
                    while self.peek()[1] == ",":
                        self.next()
                        args.append(self.parse())
                self.expect(")")
                if v in INTRINSICS and v not in self.arrays:
                    return f"{INTRINSICS[v]}({', '.join(args)})"
                return f"{v}[{', '.join(args)}]"      # array element
            return v
        raise SyntaxError(f"unexpected token {v!r}")


def tr_expr(s, arrays=()):
    e = Expr(tokenize(s), {a.lower() for a in arrays})
    out = e.parse()
    if e.p != len(e.t):
        raise SyntaxError(f"trailing tokens in {s!r}")
    return out


# ------------------------------------------------------------------------------------------
# subroutine translator
# ------------------------------------------------------------------------------------------
STUBS = {"ice_haloupdate"}
DECL = re.compile(r"^\s*(integer|real|logical|character|type)\b", re.I)


def split_top(s, sep=","):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


class Sub:
    """one reference subroutine: its interface (from the declarations) and its body as Python source."""

    def __init__(self, path, name, registry):
        self.name, self.registry = name, registry
        L = subroutine_lines(path, name)
        m = re.match(r"^\s*subroutine\s+\w+\s*\((.*)\)\s*$", L[0], re.I)
        self.args = [a.strip().lower() for a in m.group(1).split(",")]
        self.arrays, self.out_scalars, self.locals_ = set(), [], set()
        self.local_arrays = []   # (name, [dims]) declared in the body, not dummy arguments
        body = []
        for ln in L[1:-1]:
            if re.match(r"^\s*(use|implicit)\b", ln, re.I):
                continue
            if DECL.match(ln) and "::" in ln:
                attrs, names = ln.split("::", 1)
                is_arr = re.search(r"dimension\s*\(", attrs, re.I) is not None
                mdim = re.search(r"dimension\s*\(([^)]*)\)", attrs, re.I)
                intent = re.search(r"intent\s*\(\s*(\w+)\s*\)", attrs, re.I)
                if re.search(r"\bparameter\b", attrs, re.I):
                    continue
                for nm in split_top(names):
                    base = re.match(r"\s*(\w+)", nm).group(1).lower()
                    if is_arr or "(" in nm:
                        self.arrays.add(base)
                        if base not in self.args:
                            dims = mdim.group(1) if mdim else nm[nm.index("(") + 1:nm.rindex(")")]
                            self.local_arrays.append((base, split_top(dims)))
                    elif intent and intent.group(1).lower() in ("out", "inout") and base in self.args:
                        self.out_scalars.append(base)
                continue
            body.append(ln)
        self.out_scalars.sort(key=self.args.index)
        self.body = body

    def python(self):
        src = [f"def {self.name.lower()}({', '.join(self.args)}):"]
        ind = 1
        sel = []   # stack of (selector expression, first case seen) for select case

        def emit(s):
            src.append("    " * ind + s)
            if s.endswith(":"):                      # a block whose statements are all skipped must not be empty
                src.append("    " * (ind + 1) + "pass")

        if getattr(self, "module_vars", None):
            emit("global " + ", ".join(self.module_vars))     # module variables the routine assigns
        for nm, dims in self.local_arrays:
            emit(f"{nm} = _alloc({', '.join(tr_expr(d, self.arrays) for d in dims)})")

        for ln in self.body:
            low = ln.lower().strip()
            if "icepack_warnings" in low or re.match(r"^write\s*\(", low):
                continue                                   # diagnostics plumbing, not arithmetic
            ln = re.sub(r",\s*kind\s*=\s*\w+\s*\)", ")", ln)   # real(ndte,kind=dbl_kind) -> real(ndte)
            m = re.match(r"^call\s+icepack_query_parameters\s*\((.*)\)$", ln.strip(), re.I)
            if m:
                for kw in split_top(m.group(1)):
                    mm = re.match(r"^(\w+)_out\s*=\s*(\w+)$", kw.strip(), re.I)
                    emit(f"{mm.group(2).lower()} = ICEPACK[{mm.group(1).lower()!r}]")
                continue
            m = re.match(r"^do\s+(\w+)\s*=\s*(.+)$", ln.strip(), re.I)
            if m:
                lo, hi = split_top(m.group(2))[:2]
                emit(f"for {m.group(1).lower()} in range({tr_expr(lo, self.arrays)}, ({tr_expr(hi, self.arrays)}) + 1):")
                ind += 1
                continue
            if re.match(r"^end\s*do\b", low):
                ind -= 1
                continue
            if re.match(r"^call\s+abort_ice\b", low):
                emit(f"raise RuntimeError({ln.strip()!r})")
                continue
            m = re.match(r"^this_block\s*=\s*get_block\s*\(", ln.strip(), re.I)
            if m:
                emit("this_block = get_block(iblk)")      # ice_blocks.F90 get_block: the block table entry
                continue
            m = re.match(r"^select\s+case\s*\((.*)\)$", ln.strip(), re.I)
            if m:
                sel.append([tr_expr(m.group(1), self.arrays), False])
                continue
            m = re.match(r"^case\s*\((.*)\)$", ln.strip(), re.I)
            if m:
                vals = ", ".join(tr_expr(v, self.arrays) for v in split_top(m.group(1)))
                if sel[-1][1]:
                    ind -= 1
                emit(f"{'elif' if sel[-1][1] else 'if'} {sel[-1][0]} in ({vals},):")
                sel[-1][1] = True
                ind += 1
                continue
            if re.match(r"^case\s+default$", low):
                if sel[-1][1]:
                    ind -= 1
                    emit("else:")
                else:
                    emit("if True:")
                sel[-1][1] = True
                ind += 1
                continue
            if re.match(r"^end\s*select\b", low):
                if sel.pop()[1]:
                    ind -= 1
                continue
            m = re.match(r"^if\s*\((.*)\)\s*then$", ln.strip(), re.I)
            if m:
                emit(f"if {tr_expr(m.group(1), self.arrays)}:")
                ind += 1
                continue
            m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", ln.strip(), re.I)
            if m:
                ind -= 1
                emit(f"elif {tr_expr(m.group(1), self.arrays)}:")
                ind += 1
                continue
            if low == "else":
                ind -= 1
                emit("else:")
                ind += 1
                continue
            if re.match(r"^end\s*if\b", low):
                ind -= 1
                continue
            if re.match(r"^if\s*\(", ln.strip(), re.I):       # one-line  if (cond) statement
                st = ln.strip()
                depth, k = 0, st.index("(")
                for k in range(st.index("("), len(st)):
                    depth += st[k] == "("
                    depth -= st[k] == ")"
                    if depth == 0:
                        break
                emit(f"if {tr_expr(st[st.index('(') + 1:k], self.arrays)}:")
                ind += 1
                ln = st[k + 1:].strip()
                one_line = True
            else:
                one_line = False
            m = re.match(r"^call\s+(\w+)\s*\((.*)\)$", ln.strip(), re.I)
            if m:
                actual = split_top(m.group(2))
                if m.group(1).lower() in STUBS:       # e.g. the MPI halo: done by the driver, not by the transliterated text
                    emit("pass")
                    if one_line:
                        ind -= 1
                    continue
                if m.group(1) not in self.registry:   # a routine this script does not transliterate (never reached by the driver)
                    emit(f"raise RuntimeError('{m.group(1)} is not transliterated')")
                    if one_line:
                        ind -= 1
                    continue
                callee = self.registry[m.group(1)]
                if isinstance(callee, list):      # generic interface: resolved by the number of arguments
                    callee = [c for c in callee if len(c.args) == len(actual)][0]
                if len(actual) != len(callee.args):
                    raise SyntaxError(f"{self.name}: call {m.group(1)} with {len(actual)} args, expected {len(callee.args)}")
                outs = [tr_expr(actual[callee.args.index(o)], self.arrays) for o in callee.out_scalars]
                pyargs = ["None" if callee.args[k] in callee.out_scalars and re.fullmatch(r"\w+", a) else tr_expr(a, self.arrays)
                          for k, a in enumerate(actual)]
                lhs = f"{', '.join(outs)}{',' if len(outs) == 1 else ''} = " if outs else ""
                emit(f"{lhs}{callee.name.lower()}({', '.join(pyargs)})")
                if one_line:
                    ind -= 1
                continue
            m = re.match(r"^(\w+)\s*\(\s*:(\s*,\s*:)*\s*\)\s*=\s*(.+)$", ln.strip())
            if m:                                          # whole-array assignment  a(:,:,:) = c0
                emit(f"{m.group(1).lower()}.fill({tr_expr(m.group(3), self.arrays)})")
                continue
            # assignment: split at the first top-level '=' that is not part of a relational operator
            depth, eq = 0, None
            for k, ch in enumerate(ln):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "=" and depth == 0 and ln[k + 1:k + 2] != "=" and ln[k - 1:k] not in ("=", "/", "<", ">"):
                    eq = k
                    break
            if eq is None:
                raise SyntaxError(f"{self.name}: cannot translate statement {ln!r}")
            emit(f"{tr_expr(ln[:eq], self.arrays)} = {tr_expr(ln[eq + 1:], self.arrays)}")
            if one_line:
                ind -= 1
        ret = ", ".join(self.out_scalars)
        src.append(f"    return ({ret}{',' if len(self.out_scalars) == 1 else ''})" if self.out_scalars else "    return None")
        return "\n".join(src)


def reference_constants():
    """evaluate the `real (kind=dbl_kind), parameter` definitions of ice_constants.F90 in source order."""
    env = {}
    for ln in logical_lines(open(os.path.join(REF, F_CONST)).read()):
        if not (re.match(r"^\s*real\b", ln, re.I) and re.search(r"\bparameter\b", ln, re.I) and "::" in ln):
            continue
        for item in split_top(ln.split("::", 1)[1]):
            if "=" not in item:
                continue
            nm, ex = item.split("=", 1)
            try:
                env[nm.strip().lower()] = float(eval(tr_expr(ex), {"math": math, "_sq": lambda x: x * x}, dict(env)))
            except Exception:
                pass  # parameters built from names defined elsewhere (not needed by the dynamics)
    return env


# ------------------------------------------------------------------------------------------
# Fortran arrays: 1-based, first index fastest, over numpy arrays stored (..., ny, nx)
# ------------------------------------------------------------------------------------------
class FArr:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            return self.a[idx - 1].item()
        return self.a[tuple(k - 1 for k in reversed(idx))].item()

    def __setitem__(self, idx, v):
        if not isinstance(idx, tuple):
            self.a[idx - 1] = v
        else:
            self.a[tuple(k - 1 for k in reversed(idx))] = v

    def fill(self, v):
        self.a[...] = v


def build_reference_functions(params):
    reg = {}
    reg["strain_rates"] = Sub(F_SHARED, "strain_rates", reg)
    reg["visc_replpress"] = Sub(F_SHARED, "visc_replpress", reg)
    reg["stress"] = Sub(F_EVP, "stress", reg)
    reg["stepu"] = Sub(F_SHARED, "stepu", reg)
    env = {"math": math, "_sq": lambda x: x * x, "_sign": lambda a, b: math.copysign(abs(a), b), "_trim": lambda s: s.strip(),
           "ICEPACK": {"rhow": params["rhow"]}}
    env.update(reference_constants())
    # module variables of ice_dyn_shared the routines read (set_evp_parameters, ice_dyn_shared.F90:453-486)
    for k in ("arlx1i", "denom1", "revp", "brlx", "e_factor", "epp2i", "capping", "Ktens", "u0", "cosw", "sinw"):
        env[k] = float(params[k])
    env = {k.lower() if k not in ("ICEPACK",) else k: v for k, v in env.items()}
    srcs = {}
    for nm in ("strain_rates", "visc_replpress", "stress", "stepu"):
        srcs[nm] = reg[nm].python()
        exec(compile(srcs[nm], f"<{nm} transliterated from {REF}>", "exec"), env)
    return env, srcs


# ------------------------------------------------------------------------------------------
# the subcycle loop of ice_dyn_evp.F90:859-913 on ONE block, driven with the reference's index lists
# ------------------------------------------------------------------------------------------
def run_reference_loop(case, ndte):
    """case: cice_b200.synth.Case with a single block (E-W cyclic, N-S closed/open).  Returns the inout fields."""
    g, p = case.grid, dict(case.params)
    assert g["nblocks"] == 1
    nxb, nyb = g["nx_block"], g["ny_block"]
    ilo, ihi, jlo, jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    env, _ = build_reference_functions(p)
    f = {k: v[0].copy() for k, v in case.fields.items()}       # (ny_block, nx_block)
    geo = {k: np.asarray(g[k][0]) for k in ("dxT", "dyT", "dxhy", "dyhx", "cxp", "cyp", "cxm", "cym", "DminTarea", "uarear")}
    A = {k: FArr(v) for k, v in {**f, **geo}.items()}
    # index lists exactly as dyn_prep2 builds them (ice_dyn_shared.F90:740-749, 759-770)
    Ti, Tj, Ui, Uj = [], [], [], []
    for j in range(jlo, jhi + 2):
        for i in range(ilo, ihi + 2):
            if f["iceTmask"][j - 1, i - 1]:
                Ti.append(i); Tj.append(j)
    for j in range(jlo, jhi + 1):
        for i in range(ilo, ihi + 1):
            if f["iceUmask"][j - 1, i - 1]:
                Ui.append(i); Uj.append(j)
    indxTi, indxTj = FArr(np.array(Ti or [0])), FArr(np.array(Tj or [0]))
    indxUi, indxUj = FArr(np.array(Ui or [0])), FArr(np.array(Uj or [0]))
    strtmp = FArr(np.zeros((8, nyb, nxb)))
    uinit, vinit = FArr(f["uvel"].copy()), FArr(f["vvel"].copy())
    cyclic = g["ew_boundary_type"] == 2
    for _ in range(ndte):
        env["stress"](nxb, nyb, len(Ti), indxTi, indxTj, A["uvel"], A["vvel"], A["dxT"], A["dyT"], A["dxhy"], A["dyhx"],
                      A["cxp"], A["cyp"], A["cxm"], A["cym"], A["DminTarea"], A["strength"],
                      A["stressp_1"], A["stressp_2"], A["stressp_3"], A["stressp_4"], A["stressm_1"], A["stressm_2"],
                      A["stressm_3"], A["stressm_4"], A["stress12_1"], A["stress12_2"], A["stress12_3"], A["stress12_4"], strtmp)
        env["stepu"](nxb, nyb, len(Ui), A["cdn_ocnU"], indxUi, indxUj, A["aiU"], strtmp, A["uocnU"], A["vocnU"],
                     A["waterxU"], A["wateryU"], A["forcexU"], A["forceyU"], A["umassdti"], A["fmU"], A["uarear"],
                     A["strintxU"], A["strintyU"], A["taubxU"], A["taubyU"], uinit, vinit, A["uvel"], A["vvel"], A["TbU"])
        # dyn_haloUpdate(uvel, vvel) for one block that spans the domain: E-W cyclic wrap, N-S closed (ghost rows untouched)
        if cyclic:
            for a in (f["uvel"], f["vvel"]):
                a[:, ilo - 2] = a[:, ihi - 1]
                a[:, ihi] = a[:, ilo - 1]
    return f


# ------------------------------------------------------------------------------------------
# C grid: the subcycle loop of ice_dyn_evp.F90:936-1101 on ONE block
# ------------------------------------------------------------------------------------------
F_GRID = "cicecore/cicedyn/infrastructure/ice_grid.F90"


class _Block:
    pass


def build_reference_cfunctions(params, case):
    reg = {}
    for path, name in ((F_SHARED, "visc_replpress"), (F_SHARED, "strain_rates_Tdt"), (F_SHARED, "strain_rates_Tdtsd"),
                       (F_SHARED, "strain_rates_U"), (F_EVP, "stressC_T"), (F_EVP, "stressC_U"), (F_EVP, "div_stress_Ex"),
                       (F_EVP, "div_stress_Ny"), (F_SHARED, "stepu_C"), (F_SHARED, "stepv_C"), (F_GRID, "grid_average_X2YS"),
                       (F_GRID, "grid_average_X2YA"), (F_GRID, "grid_average_X2Y_1")):
        reg[name] = Sub(path, name, reg)
    reg["strain_rates_T"] = [reg["strain_rates_Tdt"], reg["strain_rates_Tdtsd"]]   # interface, ice_dyn_shared.F90:146
    g, cg = case.grid, case.cgrid
    blk = _Block()
    blk.ilo, blk.ihi, blk.jlo, blk.jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    env = {"math": math, "_sq": lambda x: x * x, "_sign": lambda a, b: math.copysign(abs(a), b), "_trim": lambda s: s.strip(),
           "_alloc": lambda nx, ny: FArr(np.zeros((ny, nx))), "ICEPACK": {"rhow": params["rhow"]},
           "nblocks": 1, "get_block": lambda iblk: blk}
    env.update(reference_constants())
    for k in ("arlx1i", "denom1", "revp", "brlx", "e_factor", "epp2i", "capping", "Ktens", "u0", "cosw", "sinw", "deltaminEVP"):
        env[k] = float(params[k])
    env["visc_method"] = "avg_strength" if params["visc_method"] == 1 else "avg_zeta"   # dynamics_nml visc_method (ice_init.F90:453)
    for k in ("tarea", "hm", "uvm", "earea", "narea", "epm", "npm", "uarea"):            # ice_grid module arrays of grid_average_X2Y_1
        env[k] = FArr(np.asarray(cg[k]))
    env = {k.lower() if k not in ("ICEPACK",) else k: v for k, v in env.items()}
    for nm in ("visc_replpress", "strain_rates_Tdt", "strain_rates_Tdtsd", "strain_rates_U", "stressC_T", "stressC_U", "div_stress_Ex",
               "div_stress_Ny", "stepu_C", "stepv_C", "grid_average_X2YS", "grid_average_X2YA", "grid_average_X2Y_1"):
        exec(compile(reg[nm].python(), f"<{nm} transliterated from {REF}>", "exec"), env)
    return env


def run_reference_cloop(case, ndte, halo_update):
    """case: cice_b200.synth.CCase with a single block.  halo_update(arrays, field_loc, field_type): the ghost-cell refresh
    (the halo is not what is being pinned here; the oracle's own halo, pinned by halochk's closed form, is passed in)."""
    g, cg, p = case.grid, case.cgrid, dict(case.params)
    assert g["nblocks"] == 1
    nxb, nyb = g["nx_block"], g["ny_block"]
    ilo, ihi, jlo, jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    env = build_reference_cfunctions(p, case)
    f = {k: np.ascontiguousarray(v.copy()) for k, v in case.fields.items()}     # (1, ny_block, nx_block)
    A3 = {k: FArr(v) for k, v in f.items()}                                     # (i,j,iblk) views for grid_average
    A = {k: FArr(v[0]) for k, v in f.items()}
    G = {k: FArr(np.asarray(cg[k])[0]) for k in cg}
    for k in ("dxT", "dyT", "DminTarea"):
        G[k] = FArr(np.asarray(g[k])[0])

    def index_list(mask, j1, i1):
        I, J = [], []
        for j in range(jlo, j1 + 1):
            for i in range(ilo, i1 + 1):
                if mask[0, j - 1, i - 1]:
                    I.append(i); J.append(j)
        return len(I), FArr(np.array(I or [0])), FArr(np.array(J or [0]))

    nT, Ti, Tj = index_list(f["iceTmask"], jhi + 1, ihi + 1)     # T list includes the N/E ghost cells (shared.F90:740-749)
    nU, Ui, Uj = index_list(f["iceUmask"], jhi, ihi)
    nE, Ei, Ej = index_list(f["iceEmask"], jhi, ihi)
    nN, Ni, Nj = index_list(f["iceNmask"], jhi, ihi)
    uE0, vN0 = FArr(f["uvelE"][0].copy()), FArr(f["vvelN"][0].copy())
    CENTER, NE, SCALAR, VECTOR = 0, 1, 0, 1
    X2Y = env["grid_average_x2y_1"]       # grid_average_X2Y_base builds X2Y = grid1//'2'//grid2//type (ice_grid.F90:3817-3841)
    for _ in range(ndte):                 # ice_dyn_evp.F90:938
        env["strain_rates_u"](nxb, nyb, nU, Ui, Uj, A["uvelE"], A["vvelE"], A["uvelN"], A["vvelN"], A["uvel"], A["vvel"], G["dxE"], G["dyN"],
                              G["dxU"], G["dyU"], G["ratiodxN"], G["ratiodxNr"], G["ratiodyE"], G["ratiodyEr"], G["epm"], G["npm"],
                              A["divergU"], A["tensionU"], A["shearU"], A["deltaU"])
        halo_update([f["shearU"]], NE, SCALAR)
        env["stressc_t"](nxb, nyb, nT, Ti, Tj, A["uvelE"], A["vvelE"], A["uvelN"], A["vvelN"], G["dxN"], G["dyE"], G["dxT"], G["dyT"],
                         G["uarea"], G["DminTarea"], A["strength"], A["shearU"], A["zetax2T"], A["etax2T"], A["stresspT"], A["stressmT"],
                         A["stress12T"])
        halo_update([f["zetax2T"], f["etax2T"], f["stresspT"], f["stressmT"]], CENTER, SCALAR)
        if env["visc_method"] == "avg_strength":
            X2Y("T2US", A3["strength"], A3["strengthU"])
        else:
            X2Y("T2US", A3["etax2T"], A3["etax2U"])
        env["stressc_u"](nxb, nyb, nU, Ui, Uj, G["uarea"], A["etax2U"], A["deltaU"], A["strengthU"], A["shearU"], A["stress12U"])
        halo_update([f["stress12U"]], NE, SCALAR)
        env["div_stress_ex"](nxb, nyb, nE, Ei, Ej, G["dxE"], G["dyE"], G["dxU"], G["dyT"], G["earear"], A["rheofactE"], A["stresspT"],
                             A["stressmT"], A["stress12U"], A["strintxE"])
        env["div_stress_ny"](nxb, nyb, nN, Ni, Nj, G["dxN"], G["dyN"], G["dxT"], G["dyU"], G["narear"], A["rheofactN"], A["stresspT"],
                             A["stressmT"], A["stress12U"], A["strintyN"])
        env["stepu_c"](nxb, nyb, nE, A["cdn_ocnE"], Ei, Ej, A["aiE"], A["uocnE"], A["vocnE"], A["waterxE"], A["forcexE"], A["emassdti"],
                       A["fmE"], A["strintxE"], A["taubxE"], uE0, A["uvelE"], A["vvelE"], A["TbE"])
        env["stepv_c"](nxb, nyb, nN, A["cdn_ocnN"], Ni, Nj, A["aiN"], A["uocnN"], A["vocnN"], A["wateryN"], A["forceyN"], A["nmassdti"],
                       A["fmN"], A["strintyN"], A["taubyN"], vN0, A["uvelN"], A["vvelN"], A["TbN"])
        halo_update([f["uvelE"]], NE, VECTOR)
        halo_update([f["vvelN"]], NE, VECTOR)
        X2Y("E2NA", A3["uvelE"], A3["uvelN"])
        X2Y("N2EA", A3["vvelN"], A3["vvelE"])
        f["uvelN"][...] = f["uvelN"] * np.asarray(cg["npm"])      # uvelN(:,:,:) = uvelN(:,:,:)*npm(:,:,:)   evp.F90:1075
        f["vvelE"][...] = f["vvelE"] * np.asarray(cg["epm"])
        halo_update([f["uvelN"]], NE, VECTOR)
        halo_update([f["vvelE"]], NE, VECTOR)
        X2Y("E2UA", A3["uvelE"], A3["uvel"])
        X2Y("N2UA", A3["vvelN"], A3["vvel"])
        f["uvel"][...] = f["uvel"] * np.asarray(cg["uvm"])
        f["vvel"][...] = f["vvel"] * np.asarray(cg["uvm"])
        halo_update([f["uvel"], f["vvel"]], NE, VECTOR)
    return f


def run_reference_cdloop(case, ndte, halo_update):
    """grid_ice = 'CD': the subcycle loop of ice_dyn_evp.F90:1123-1275 on ONE block (same conventions as run_reference_cloop)."""
    g, cg, p = case.grid, case.cgrid, dict(case.params)
    assert g["nblocks"] == 1
    nxb, nyb = g["nx_block"], g["ny_block"]
    ilo, ihi, jlo, jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    reg = {}
    names = ((F_SHARED, "visc_replpress"), (F_SHARED, "strain_rates_Tdt"), (F_SHARED, "strain_rates_Tdtsd"), (F_SHARED, "strain_rates_U"),
             (F_EVP, "stressCD_T"), (F_EVP, "stressCD_U"), (F_EVP, "div_stress_Ex"), (F_EVP, "div_stress_Ey"), (F_EVP, "div_stress_Nx"),
             (F_EVP, "div_stress_Ny"), (F_SHARED, "stepuv_CD"), (F_GRID, "grid_average_X2YS"), (F_GRID, "grid_average_X2YA"),
             (F_GRID, "grid_average_X2Y_1"))
    for path, name in names:
        reg[name] = Sub(path, name, reg)
    reg["strain_rates_T"] = [reg["strain_rates_Tdt"], reg["strain_rates_Tdtsd"]]
    blk = _Block()
    blk.ilo, blk.ihi, blk.jlo, blk.jhi = ilo, ihi, jlo, jhi
    env = {"math": math, "_sq": lambda x: x * x, "_sign": lambda a, b: math.copysign(abs(a), b), "_trim": lambda s: s.strip(),
           "_alloc": lambda nx, ny: FArr(np.zeros((ny, nx))), "ICEPACK": {"rhow": p["rhow"]}, "nblocks": 1, "get_block": lambda iblk: blk}
    env.update(reference_constants())
    for k in ("arlx1i", "denom1", "revp", "brlx", "e_factor", "epp2i", "capping", "Ktens", "u0", "cosw", "sinw", "deltaminEVP"):
        env[k] = float(p[k])
    env["visc_method"] = "avg_strength" if p["visc_method"] == 1 else "avg_zeta"
    for k in ("tarea", "hm", "uvm", "earea", "narea", "epm", "npm", "uarea"):
        env[k] = FArr(np.asarray(cg[k]))
    env = {k.lower() if k != "ICEPACK" else k: v for k, v in env.items()}
    for _, nm in names:
        exec(compile(reg[nm].python(), f"<{nm} transliterated from {REF}>", "exec"), env)
    f = {k: np.ascontiguousarray(v.copy()) for k, v in case.fields.items()}
    A3 = {k: FArr(v) for k, v in f.items()}
    A = {k: FArr(v[0]) for k, v in f.items()}
    G = {k: FArr(np.asarray(cg[k])[0]) for k in cg}
    for k in ("dxT", "dyT", "DminTarea"):
        G[k] = FArr(np.asarray(g[k])[0])

    def index_list(mask, j1, i1):
        I, J = [], []
        for j in range(jlo, j1 + 1):
            for i in range(ilo, i1 + 1):
                if mask[0, j - 1, i - 1]:
                    I.append(i); J.append(j)
        return len(I), FArr(np.array(I or [0])), FArr(np.array(J or [0]))

    nT, Ti, Tj = index_list(f["iceTmask"], jhi + 1, ihi + 1)
    nU, Ui, Uj = index_list(f["iceUmask"], jhi, ihi)
    nE, Ei, Ej = index_list(f["iceEmask"], jhi, ihi)
    nN, Ni, Nj = index_list(f["iceNmask"], jhi, ihi)
    init = {k: FArr(f[k][0].copy()) for k in ("uvelE", "vvelE", "uvelN", "vvelN")}
    CENTER, NE, SCALAR, VECTOR = 0, 1, 0, 1
    X2Y = env["grid_average_x2y_1"]
    for _ in range(ndte):                 # ice_dyn_evp.F90:1125
        env["stresscd_t"](nxb, nyb, nT, Ti, Tj, A["uvelE"], A["vvelE"], A["uvelN"], A["vvelN"], G["dxN"], G["dyE"], G["dxT"], G["dyT"],
                          G["DminTarea"], A["strength"], A["zetax2T"], A["etax2T"], A["stresspT"], A["stressmT"], A["stress12T"])
        halo_update([f["zetax2T"], f["etax2T"]], CENTER, SCALAR)
        if env["visc_method"] == "avg_strength":
            X2Y("T2US", A3["strength"], A3["strengthU"])
        else:
            X2Y("T2US", A3["zetax2T"], A3["zetax2U"])
            X2Y("T2US", A3["etax2T"], A3["etax2U"])
        env["strain_rates_u"](nxb, nyb, nU, Ui, Uj, A["uvelE"], A["vvelE"], A["uvelN"], A["vvelN"], A["uvel"], A["vvel"], G["dxE"], G["dyN"],
                              G["dxU"], G["dyU"], G["ratiodxN"], G["ratiodxNr"], G["ratiodyE"], G["ratiodyEr"], G["epm"], G["npm"],
                              A["divergU"], A["tensionU"], A["shearU"], A["deltaU"])
        env["stresscd_u"](nxb, nyb, nU, Ui, Uj, G["uarea"], A["zetax2U"], A["etax2U"], A["strengthU"], A["divergU"], A["tensionU"],
                          A["shearU"], A["deltaU"], A["stresspU"], A["stressmU"], A["stress12U"])
        halo_update([f["stresspT"], f["stressmT"], f["stress12T"]], CENTER, SCALAR)
        halo_update([f["stresspU"], f["stressmU"], f["stress12U"]], NE, SCALAR)
        env["div_stress_ex"](nxb, nyb, nE, Ei, Ej, G["dxE"], G["dyE"], G["dxU"], G["dyT"], G["earear"], A["rheofactE"], A["stresspT"],
                             A["stressmT"], A["stress12U"], A["strintxE"])
        env["div_stress_ey"](nxb, nyb, nE, Ei, Ej, G["dxE"], G["dyE"], G["dxU"], G["dyT"], G["earear"], A["rheofactE"], A["stresspU"],
                             A["stressmU"], A["stress12T"], A["strintyE"])
        env["div_stress_nx"](nxb, nyb, nN, Ni, Nj, G["dxN"], G["dyN"], G["dxT"], G["dyU"], G["narear"], A["rheofactN"], A["stresspU"],
                             A["stressmU"], A["stress12T"], A["strintxN"])
        env["div_stress_ny"](nxb, nyb, nN, Ni, Nj, G["dxN"], G["dyN"], G["dxT"], G["dyU"], G["narear"], A["rheofactN"], A["stresspT"],
                             A["stressmT"], A["stress12U"], A["strintyN"])
        env["stepuv_cd"](nxb, nyb, nE, A["cdn_ocnE"], Ei, Ej, A["aiE"], A["uocnE"], A["vocnE"], A["waterxE"], A["wateryE"], A["forcexE"],
                         A["forceyE"], A["emassdti"], A["fmE"], A["strintxE"], A["strintyE"], A["taubxE"], A["taubyE"], init["uvelE"],
                         init["vvelE"], A["uvelE"], A["vvelE"], A["TbE"])
        env["stepuv_cd"](nxb, nyb, nN, A["cdn_ocnN"], Ni, Nj, A["aiN"], A["uocnN"], A["vocnN"], A["waterxN"], A["wateryN"], A["forcexN"],
                         A["forceyN"], A["nmassdti"], A["fmN"], A["strintxN"], A["strintyN"], A["taubxN"], A["taubyN"], init["uvelN"],
                         init["vvelN"], A["uvelN"], A["vvelN"], A["TbN"])
        halo_update([f["uvelE"], f["vvelE"]], NE, VECTOR)
        halo_update([f["uvelN"], f["vvelN"]], NE, VECTOR)
        X2Y("E2UA", A3["uvelE"], A3["uvel"])
        X2Y("N2UA", A3["vvelN"], A3["vvel"])
        f["uvel"][...] = f["uvel"] * np.asarray(cg["uvm"])
        f["vvel"][...] = f["vvel"] * np.asarray(cg["uvm"])
        halo_update([f["uvel"], f["vvel"]], NE, VECTOR)
    return f


CDCASES = [dict(config="tiny", seed=91, ndte=4), dict(config="tiny", seed=92, ndte=3, revised_evp=True),
           dict(config="tiny", seed=93, ndte=3, visc_method=1), dict(config="tiny", ndte=5),
           dict(config="tiny", seed=94, ndte=3, ew="closed", ns="closed"), dict(config="gx3", seed=95, ndte=2)]
CDFIELDS = ("uvelE", "vvelE", "uvelN", "vvelN", "uvel", "vvel", "stresspT", "stressmT", "stress12T", "stresspU", "stressmU", "stress12U",
            "zetax2T", "etax2T", "divergU", "tensionU", "shearU", "deltaU", "strintxE", "strintyE", "strintxN", "strintyN",
            "taubxE", "taubyE", "taubxN", "taubyN")


def generate_cd(only=None):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    from oracle import oracle
    out = {}
    for n, kw in enumerate(CDCASES):
        if only is not None and n not in only:
            continue
        c = synth.make_cdcase(**kw)
        f = run_reference_cdloop(c, c.params["ndte"], lambda arrs, loc, typ: oracle.halo_update(c.grid, arrs, loc, typ))
        for k in CDFIELDS:
            out[f"cdcase{n}_{k}"] = f[k]
    return out


CCASES = [dict(config="tiny", seed=71, ndte=4), dict(config="tiny", seed=72, ndte=3, revised_evp=True),
          dict(config="tiny", seed=73, ndte=3, visc_method=1), dict(config="tiny", ndte=5),
          dict(config="tiny", seed=74, ndte=3, ew="closed", ns="closed"), dict(config="gx3", seed=75, ndte=2)]
CFIELDS = ("uvelE", "vvelE", "uvelN", "vvelN", "uvel", "vvel", "stresspT", "stressmT", "stress12T", "stress12U", "zetax2T", "etax2T",
           "divergU", "tensionU", "shearU", "deltaU", "strintxE", "strintyN", "taubxE", "taubyN")


def generate_c(only=None):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    from oracle import oracle
    out = {}
    for n, kw in enumerate(CCASES):
        if only is not None and n not in only:
            continue
        c = synth.make_ccase(**kw)
        f = run_reference_cloop(c, c.params["ndte"], lambda arrs, loc, typ: oracle.halo_update(c.grid, arrs, loc, typ))
        for k in CFIELDS:
            out[f"ccase{n}_{k}"] = f[k]
    return out


# synth.make_case keywords (+ "params": overrides of the EVP scalars) -- seeds give random velocities, stresses and TbU (set S2)
CASES = [dict(config="tiny", seed=61, ndte=6),
         dict(config="tiny", seed=62, ndte=4, revised_evp=True),
         dict(config="tiny", seed=63, ndte=5, kmt="continents"),
         dict(config="tiny", seed=64, ndte=5, params=dict(capping=0.0, Ktens=0.2)),          # the general visc_replpress branch
         dict(config="tiny", seed=65, ndte=3, params=dict(cosw=0.9, sinw=0.4358898943540674)),  # turning angle in stepu
         dict(config="tiny", ndte=8),                                                          # set S1: box2001 start from rest
         dict(config="gx3", seed=66, ndte=3)]
FIELDS = ("stressp_1", "stressp_2", "stressp_3", "stressp_4", "stressm_1", "stressm_2", "stressm_3", "stressm_4",
          "stress12_1", "stress12_2", "stress12_3", "stress12_4", "strintxU", "strintyU", "taubxU", "taubyU", "uvel", "vvel")


def make(synth, kw):
    kw = dict(kw)
    over = kw.pop("params", {})
    c = synth.make_case(**kw)
    c.params.update(over)
    return c


def generate(only=None):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    out = {}
    for n, kw in enumerate(CASES):
        if only is not None and n not in only:
            continue
        c = make(synth, kw)
        f = run_reference_loop(c, c.params["ndte"])
        for k in FIELDS:
            out[f"case{n}_{k}"] = f[k]
    return out


# ------------------------------------------------------------------------------------------
# deformations (ice_dyn_shared.F90:1756-1860) on the random velocities of a set-S2 case
# ------------------------------------------------------------------------------------------
DFIELDS = ("divu", "shear", "vort", "rdg_conv", "rdg_shear")


def deform_inputs(synth, seed=81):
    c = synth.make_case("tiny", seed=seed, ndte=1)
    X = c.X
    tarear = np.where(X["tarea"] > 0, 1.0 / np.where(X["tarea"] > 0, X["tarea"], 1.0), 0.0)
    d = {n: synth.scatter(a, c.blocks) for n, a in (("dxU", X["dxU"]), ("dyU", X["dyU"]), ("tarear", tarear))}
    for n in DFIELDS:
        d[n] = np.full(c.fields["uvel"].shape, -7.0)
    return c, d


def generate_d():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    c, d = deform_inputs(synth)
    g, p = c.grid, c.params
    reg = {}
    reg["strain_rates"] = Sub(F_SHARED, "strain_rates", reg)
    reg["deformations"] = Sub(F_SHARED, "deformations", reg)
    env = {"math": math, "_sq": lambda x: x * x, "_sign": lambda a, b: math.copysign(abs(a), b), "e_factor": float(p["e_factor"])}
    env.update(reference_constants())
    for nm in ("strain_rates", "deformations"):
        exec(compile(reg[nm].python(), f"<{nm} transliterated from {REF}>", "exec"), env)
    ilo, ihi, jlo, jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    Ti, Tj = [], []
    for j in range(jlo, jhi + 2):
        for i in range(ilo, ihi + 2):
            if c.fields["iceTmask"][0, j - 1, i - 1]:
                Ti.append(i); Tj.append(j)
    A = lambda a: FArr(np.asarray(a)[0])
    out = {k: d[k].copy() for k in DFIELDS}
    env["deformations"](g["nx_block"], g["ny_block"], len(Ti), FArr(np.array(Ti)), FArr(np.array(Tj)), A(c.fields["uvel"]), A(c.fields["vvel"]),
                        A(g["dxT"]), A(g["dyT"]), A(d["dxU"]), A(d["dyU"]), A(g["cxp"]), A(g["cyp"]), A(g["cxm"]), A(g["cym"]), A(d["tarear"]),
                        FArr(out["vort"][0]), FArr(out["shear"][0]), FArr(out["divu"][0]), FArr(out["rdg_conv"][0]), FArr(out["rdg_shear"][0]))
    return {f"dcase0_{k}": v for k, v in out.items()}


# ------------------------------------------------------------------------------------------
# dyn_finish (ice_dyn_shared.F90:1291-1365): the ice-ocean stress after the loop, on the state a set-S2
# case reaches after 3 subcycles, with a turning angle so that the sinw*sign(fm) terms are exercised
# ------------------------------------------------------------------------------------------
FFIELDS = ("strocnxU", "strocnyU")
FINISH_ANGLE = dict(cosw=0.9, sinw=0.4358898943540674)


def finish_inputs(synth, oracle_mod, seed=82):
    c = synth.make_case("tiny", seed=seed, ndte=3)
    c.params.update(FINISH_ANGLE)
    f = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, f)           # the velocities dyn_finish sees are the loop's final ones
    d = {n: np.full(f["uvel"].shape, -7.0) for n in FFIELDS}
    return c, f, d


def generate_f():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    from oracle import oracle as oracle_mod
    c, f, d = finish_inputs(synth, oracle_mod)
    g, p = c.grid, c.params
    reg = {}
    reg["dyn_finish"] = Sub(F_SHARED, "dyn_finish", reg)
    env = {"math": math, "_sq": lambda x: x * x, "_sign": lambda a, b: math.copysign(abs(a), b), "ICEPACK": {"rhow": float(p["rhow"])},
           "cosw": float(p["cosw"]), "sinw": float(p["sinw"])}
    env.update(reference_constants())
    env = {k.lower() if k != "ICEPACK" else k: v for k, v in env.items()}
    exec(compile(reg["dyn_finish"].python(), f"<dyn_finish transliterated from {REF}>", "exec"), env)
    ilo, ihi, jlo, jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    Ui, Uj = [], []
    for j in range(jlo, jhi + 1):
        for i in range(ilo, ihi + 1):
            if f["iceUmask"][0, j - 1, i - 1]:
                Ui.append(i); Uj.append(j)
    A = lambda a: FArr(np.asarray(a, dtype=np.float64)[0])
    out = {k: d[k].copy() for k in FFIELDS}
    env["dyn_finish"](g["nx_block"], g["ny_block"], len(Ui), A(f["cdn_ocnU"]), FArr(np.array(Ui)), FArr(np.array(Uj)), A(f["uvel"]), A(f["vvel"]),
                      A(f["uocnU"]), A(f["vocnU"]), A(f["aiU"]), A(f["fmU"]), FArr(out["strocnxU"][0]), FArr(out["strocnyU"][0]))
    return {f"fcase0_{k}": v for k, v in out.items()}


# ------------------------------------------------------------------------------------------
# dyn_prep2 (ice_dyn_shared.F90:593-839): the routine that builds every time-varying input of the loop.  Run on the
# synthetic state, it must reproduce what cice_b200/synth.py feeds the kernels and the bench (SURVEY 8d inputs).
# ------------------------------------------------------------------------------------------
PFIELDS = ("umassdti", "fmU", "waterxU", "wateryU", "forcexU", "forceyU", "uvel", "vvel", "iceUmask")


def generate_prep2(config="tiny"):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    c = synth.make_case(config)           # set S1: the box2001 start, exactly what bench.py uses
    X, g = c.X, c.grid
    nxb, nyb = g["nx_block"], g["ny_block"]
    ilo, ihi, jlo, jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    reg = {}
    reg["dyn_prep2"] = Sub(F_SHARED, "dyn_prep2", reg)
    env = {"math": math, "_sq": lambda x: x * x, "_sign": lambda a, b: math.copysign(abs(a), b), "_trim": lambda s: s.strip(),
           "_present": lambda x: x is not None, "_alloc": lambda nx, ny: FArr(np.zeros((ny, nx))), "ICEPACK": {"gravit": 9.80616},
           "ssh_stress": "geostrophic", "dyn_area_min": 1e-11, "dyn_mass_min": 1e-10, "rheo_area_min": 1e-11, "cosw": 1.0, "sinw": 0.0}
    env.update(reference_constants())
    exec(compile(reg["dyn_prep2"].python(), f"<dyn_prep2 transliterated from {REF}>", "exec"), env)

    def ext(a):   # interior (ny,nx) array -> block array with zero ghosts
        out = np.zeros((nyb, nxb))
        out[jlo - 1:jhi, ilo - 1:ihi] = a
        return out

    Z = lambda: np.zeros((nyb, nxb))
    aiX, Xmass = X["aiU"].copy(), ext(X["umass_i"])
    fcor = np.full((nyb, nxb), synth.FCOR_CONST)
    Xmask = (X["uvm"] > 0.5).astype(np.float64)
    out = {k: Z() for k in ("Xmassdti", "fm", "strtltx", "strtlty", "strocnx", "strocny", "strintx", "strinty", "taubx", "tauby",
                            "waterx", "watery", "forcex", "forcey", "uvel_init", "vvel_init", "uvel", "vvel", "TbU", "iceXmask")}
    sig = [Z() for _ in range(12)]
    A = FArr
    n = nxb * nyb
    env["dyn_prep2"](nxb, nyb, ilo, ihi, jlo, jhi, None, None, A(np.zeros(n, int)), A(np.zeros(n, int)), A(np.zeros(n, int)), A(np.zeros(n, int)),
                     A(aiX), A(Xmass), A(out["Xmassdti"]), A(fcor), A(Xmask), A(X["uocnU"].copy()), A(X["vocnU"].copy()),
                     A(ext(X["strairxU_i"])), A(ext(X["strairyU_i"])), A(Z()), A(Z()), A(X["iceTmask"].astype(np.float64)), A(out["iceXmask"]),
                     A(out["fm"]), 3600.0, A(out["strtltx"]), A(out["strtlty"]), A(out["strocnx"]), A(out["strocny"]), A(out["strintx"]),
                     A(out["strinty"]), A(out["taubx"]), A(out["tauby"]), A(out["waterx"]), A(out["watery"]), A(out["forcex"]), A(out["forcey"]),
                     *[A(a) for a in sig], A(out["uvel_init"]), A(out["vvel_init"]), A(out["uvel"]), A(out["vvel"]), A(out["TbU"]), None)
    names = {"umassdti": "Xmassdti", "fmU": "fm", "waterxU": "waterx", "wateryU": "watery", "forcexU": "forcex", "forceyU": "forcey",
             "uvel": "uvel", "vvel": "vvel", "iceUmask": "iceXmask"}
    inner = (slice(jlo - 1, jhi), slice(ilo - 1, ihi))
    ref = {k: out[v][inner].copy() for k, v in names.items()}
    mine = {k: np.asarray(X[k], dtype=np.float64)[inner] for k in names}
    return ref, mine


def generate_prep1_and_averages(config="tiny"):
    """dyn_prep1 (ice_dyn_shared.F90:496-592) and the T->U averages evp() takes before dyn_prep2 (ice_dyn_evp.F90:433-456:
    grid_average_X2Y 'S' for aice, tmass, uocn, vocn; 'F' for the wind stress), run on the synthetic box2001 state."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    c = synth.make_case(config)
    X, g = c.X, c.grid
    nxb, nyb = g["nx_block"], g["ny_block"]
    blk = _Block()
    blk.ilo, blk.ihi, blk.jlo, blk.jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    reg = {}
    for path, name in ((F_SHARED, "dyn_prep1"), (F_GRID, "grid_average_X2YS"), (F_GRID, "grid_average_X2YF")):
        reg[name] = Sub(path, name, reg)
    env = {"math": math, "_sq": lambda x: x * x, "_trim": lambda s: s.strip(), "_alloc": lambda nx, ny: FArr(np.zeros((ny, nx))),
           "ICEPACK": {"rhoi": synth.RHOI, "rhos": synth.RHOS}, "dyn_area_min": 1e-11, "dyn_mass_min": 1e-10, "nblocks": 1,
           "get_block": lambda iblk: blk}
    env.update(reference_constants())
    env = {k.lower() if k != "ICEPACK" else k: v for k, v in env.items()}
    for nm in ("dyn_prep1", "grid_average_X2YS", "grid_average_X2YF"):
        exec(compile(reg[nm].python(), f"<{nm} transliterated from {REF}>", "exec"), env)
    tmass, iceT = np.zeros((nyb, nxb)), np.zeros((nyb, nxb))
    env["dyn_prep1"](nxb, nyb, blk.ilo, blk.ihi, blk.jlo, blk.jhi, FArr(X["aice"].copy()), FArr(X["vice"].copy()), FArr(np.zeros((nyb, nxb))),
                     FArr((X["hm"] > 0.5).astype(np.float64)), FArr(tmass), FArr(iceT))
    inner = (slice(blk.jlo - 1, blk.jhi), slice(blk.ilo - 1, blk.ihi))
    ref = {"tmass": tmass, "iceTmask_interior": iceT[inner].copy()}
    mine = {"tmass": X["tmass"], "iceTmask_interior": X["iceTmask"].astype(np.float64)[inner]}
    A3 = lambda a: FArr(np.ascontiguousarray(np.asarray(a, dtype=np.float64))[None, :, :])
    for out, src in (("aiU", "aice"), ("umass_i", "tmass"), ("uocnU", "uocn"), ("vocnU", "vocn")):
        w2 = np.zeros((1, nyb, nxb))
        env["grid_average_x2ys"]("NE", A3(X[src]), A3(X["tarea"]), A3(X["hm"]), FArr(w2))        # T2US, ice_grid.F90:3984
        ref[out] = w2[0][inner].copy()
        mine[out] = np.asarray(X[out], dtype=np.float64)[inner] if X[out].shape == (nyb, nxb) else np.asarray(X[out])
    for out, src in (("strairxU_i", "strax"), ("strairyU_i", "stray")):
        w2 = np.zeros((1, nyb, nxb))
        env["grid_average_x2yf"]("NE", A3(X[src]), A3(X["tarea"]), FArr(w2), A3(X["uarea"]))       # T2UF, ice_grid.F90:3958
        ref[out] = w2[0][inner].copy()
        mine[out] = np.asarray(X[out])
    return ref, mine


# ------------------------------------------------------------------------------------------
# static geometry of the B-grid loop: the tail of init_dyn_shared (ice_dyn_shared.F90), from the
# `grid_ice == 'B'` block that fills dxhy, dyhx to the loops that fill cyp, cxp, cym, cxm
# ------------------------------------------------------------------------------------------
def generate_geometry(config="tiny"):
    import tempfile
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cice_b200 import synth
    global REF
    lines = open(os.path.join(REF, F_SHARED)).read().splitlines()
    a = next(n for n, ln in enumerate(lines) if "grid_ice == 'B' .and. evp_algorithm" in ln and "if" in ln.lower()
             and n > next(k for k, l2 in enumerate(lines) if re.match(r"^\s*subroutine\s+init_dyn_shared\b", l2, re.I)))
    b = next(n for n, ln in enumerate(lines) if re.match(r"^\s*end\s+subroutine\s+init_dyn_shared\b", ln, re.I))
    text = "      subroutine geom_tail ()\n" + "\n".join(lines[a:b]) + "\n      end subroutine geom_tail\n"
    c = synth.make_case(config)
    X, g = c.X, c.grid
    blk = _Block()
    blk.ilo, blk.ihi, blk.jlo, blk.jhi = (int(g[k][0]) for k in ("ilo", "ihi", "jlo", "jhi"))
    shape = (1,) + X["HTE"].shape
    out = {k: np.zeros(shape) for k in ("dxhy", "dyhx", "cyp", "cxp", "cym", "cxm")}
    env = {"grid_ice": "B", "evp_algorithm": "standard_2d", "nblocks": 1, "get_block": lambda iblk: blk,
           "hte": FArr(X["HTE"][None].copy()), "htn": FArr(X["HTN"][None].copy()), **{k: FArr(v) for k, v in out.items()}}
    env.update(reference_constants())
    env = {k.lower(): v for k, v in env.items()}
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "g.F90"), "w").write(text)
        keep, REF = REF, tmp
        try:
            sub = Sub("g.F90", "geom_tail", {})
        finally:
            REF = keep
    sub.arrays |= {"hte", "htn", "dxhy", "dyhx", "cyp", "cxp", "cym", "cxm"}
    exec(compile(sub.python(), "<tail of init_dyn_shared transliterated>", "exec"), env)
    env["geom_tail"]()
    i0_, i1_, j0_, j1_ = blk.ilo - 1, blk.ihi, blk.jlo - 1, blk.jhi
    ref = {"dxhy": out["dxhy"][0][j0_:j1_, i0_:i1_], "dyhx": out["dyhx"][0][j0_:j1_, i0_:i1_]}
    mine = {"dxhy": X["dxhy"][j0_:j1_, i0_:i1_], "dyhx": X["dyhx"][j0_:j1_, i0_:i1_]}
    for k in ("cyp", "cxp", "cym", "cxm"):          # filled on ilo:ihi+1, jlo:jhi+1
        ref[k] = out[k][0][j0_:j1_ + 1, i0_:i1_ + 1]
        mine[k] = X[k][j0_:j1_ + 1, i0_:i1_ + 1]
    return ref, mine


# ------------------------------------------------------------------------------------------
# EVP constants: set_evp_parameters (ice_dyn_shared.F90:453-486)
# ------------------------------------------------------------------------------------------
def generate_params(ndte, revised_evp, elasticDamp=0.36, e_yieldcurve=2.0, e_plasticpot=2.0, arlx=300.0, brlx=300.0):
    reg = {}
    sub = Sub(F_SHARED, "set_evp_parameters", reg)
    sub.module_vars = ["dtei", "epp2i", "e_factor", "ecci", "revp", "denom1", "arlx1i", "arlx", "brlx"]
    env = {"ndte": int(ndte), "revised_evp": bool(revised_evp), "elasticdamp": elasticDamp, "e_yieldcurve": e_yieldcurve,
           "e_plasticpot": e_plasticpot, "arlx": arlx, "brlx": brlx, "my_task": 1, "master_task": 0, "_sq": lambda x: x * x}
    env.update({k.lower(): v for k, v in reference_constants().items()})
    exec(compile(sub.python(), "<set_evp_parameters transliterated>", "exec"), env)
    env["set_evp_parameters"](3600.0)
    return {k: float(env[k]) for k in ("epp2i", "e_factor", "revp", "denom1", "arlx1i", "brlx")}


FULL_VECTORS = (0, 1, 3)   # cases whose arrays are committed in full; every case is committed as sha256 per field
FULL_CVECTORS = (0, 2)
FULL_CDVECTORS = (0, 2)
SHA = os.path.join(HERE, "ref_source_vectors.json")


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


if __name__ == "__main__":
    import json
    vec = generate()
    cvec = generate_c()
    dvec = generate_d()
    dvec.update(generate_f())
    cdvec = generate_cd()
    pvec = {}
    for cfg in ("tiny", "gx3", "gx1"):
        pref, _ = generate_prep2(cfg)
        pvec.update({f"prep2_{cfg}_{k}": v for k, v in pref.items()})
        qref, _ = generate_prep1_and_averages(cfg)
        pvec.update({f"prep1_{cfg}_{k}": np.asarray(v, dtype=np.float64) for k, v in qref.items()})
        gref, _ = generate_geometry(cfg)
        pvec.update({f"geom_{cfg}_{k}": np.ascontiguousarray(v) for k, v in gref.items()})
    if "--write" in sys.argv:
        full = {k: v for k, v in vec.items() if int(k[4:k.index("_")]) in FULL_VECTORS}
        full.update({k: v for k, v in cvec.items() if int(k[5:k.index("_")]) in FULL_CVECTORS})
        full.update(dvec)
        full.update({k: v for k, v in cdvec.items() if int(k[6:k.index("_")]) in FULL_CDVECTORS})
        np.savez_compressed(OUT, **full)
        meta = {"_how": "python tests/golden/ref_translit.py --write  (transliterates the reference's Fortran subroutines under /root/reference "
                        "-- B grid: stress, stepu, strain_rates, visc_replpress; C grid: strain_rates_U, strain_rates_Tdt, stressC_T, stressC_U, "
                        "div_stress_Ex/Ny, stepu_C, stepv_C, grid_average_X2Y_1/X2YS/X2YA; CD grid: stressCD_T, stressCD_U, strain_rates_Tdtsd, "
                        "div_stress_Ey/Nx, stepuv_CD; deformations; dyn_finish; ice_constants -- and runs them; see the module docstring)",
                "cases": [dict(kw) for kw in CASES], "ccases": [dict(kw) for kw in CCASES],
                "cdcases": [dict(kw) for kw in CDCASES],
                "sha256": {k: sha(v) for k, v in {**vec, **cvec, **dvec, **cdvec, **pvec}.items()}}
        json.dump(meta, open(SHA, "w"), indent=1)
        print("wrote", OUT, os.path.getsize(OUT), "bytes;", SHA)
    if "--show" in sys.argv:
        _, srcs = build_reference_functions(dict(arlx1i=0, denom1=0, revp=0, brlx=0, e_factor=0, epp2i=0, capping=1, Ktens=0, u0=0, cosw=1,
                                                 sinw=0, rhow=1026))
        for k, v in srcs.items():
            print(v, "\n")
    print({k: float(np.abs(v).max()) for k, v in vec.items() if k.startswith("case0_") and k.endswith(("uvel", "stressp_1"))})
