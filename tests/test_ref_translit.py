"""Unit tests of the Fortran-subset transliterator (tests/golden/ref_translit.py) that produces the reference-source vectors:
the tool must read Fortran the way a Fortran compiler does (operator precedence and associativity, unary minus, `**2`,
continuations, comments, control flow, intent(out) scalars), otherwise the vectors would not be the reference's."""
import math
import os
import sys
import textwrap

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_translit as rt  # noqa: E402

ENV = {"math": math, "_sq": lambda x: x * x, "_sign": lambda a, b: math.copysign(abs(a), b), "_trim": lambda s: s.strip(),
       "_present": lambda x: x is not None}


def ev(expr, **names):
    return eval(rt.tr_expr(expr, arrays=[k for k, v in names.items() if isinstance(v, rt.FArr)]), dict(ENV), names)


@pytest.mark.parametrize("expr,want", [
    ("a - b - c", (7.0 - 2.0) - 0.5),                     # left to right at equal precedence
    ("a / b * c", (7.0 / 2.0) * 0.5),
    ("a - b * c", 7.0 - (2.0 * 0.5)),
    ("-a*b", -(7.0 * 2.0)),                               # unary minus binds looser than *
    ("-a**2", -(7.0 * 7.0)),                              # ... and looser than **
    ("a**2*b", (7.0 * 7.0) * 2.0),
    ("a*(b+c)**2", 7.0 * ((2.0 + 0.5) * (2.0 + 0.5))),
    ("a + -b", None),                                     # two consecutive operators: not Fortran
    ("1.5_dbl_kind*a + 2.0d0*b", 1.5 * 7.0 + 2.0 * 2.0),
    ("max(a, b*c, 3.0)", 7.0),
    ("sign(c, -b)", -0.5),
    ("sqrt(a**2 + b**2)", math.sqrt(7.0 * 7.0 + 2.0 * 2.0)),
])
def test_expression_semantics(expr, want):
    if want is None:
        with pytest.raises(SyntaxError):
            ev(expr, a=7.0, b=2.0, c=0.5)
        return
    got = ev(expr, a=7.0, b=2.0, c=0.5)
    assert got == want and math.copysign(1.0, got) == math.copysign(1.0, want)


def test_square_is_a_product_not_pow():
    x = 1.0000000000000002
    assert ev("x**2", x=x) == x * x
    assert "_sq(" in rt.tr_expr("(a - b)**2")


@pytest.mark.parametrize("expr,want", [
    (".not. p .and. q .or. r", ((not True) and False) or True),
    ("a > b .and. b >= c .or. .false.", True),
    ("a /= b", True), ("a == b", False), ("a .le. b", False),
    ("trim(s) == 'abc'", True),
])
def test_logical_and_relational(expr, want):
    assert bool(ev(expr, a=7.0, b=2.0, c=0.5, p=True, q=False, r=True, s="abc  ")) is want


def test_arrays_are_one_based_first_index_fastest():
    a = rt.FArr(np.arange(12.0).reshape(3, 4))        # Fortran a(4,3): a(i,j) = a_np[j-1, i-1]
    assert ev("a(1,1)", a=a) == 0.0 and ev("a(4,1)", a=a) == 3.0 and ev("a(1,2)", a=a) == 4.0 and ev("a(i+1,j-1)", a=a, i=2, j=3) == 6.0
    b = rt.FArr(np.arange(24.0).reshape(2, 3, 4))     # b(4,3,2)
    assert ev("b(2,3,2)", b=b) == 1 * 12 + 2 * 4 + 1
    a[2, 3] = -1.0
    assert a.a[2, 1] == -1.0


FORTRAN = textwrap.dedent("""
      subroutine inner (x, y, s, p)
      real (kind=dbl_kind), intent(in) :: x, y
      real (kind=dbl_kind), intent(out) :: &
         s , & ! a comment with an ! inside 'and a quote'
         p
      s = x + y        ! trailing comment
      p = x * &
          y
      end subroutine inner

      subroutine outer (n, a, total, mode)
      integer (kind=int_kind), intent(in) :: n
      real (kind=dbl_kind), dimension (n), intent(inout) :: a
      real (kind=dbl_kind), intent(out) :: total
      character(len=*), intent(in) :: mode
      integer (kind=int_kind) :: i
      real (kind=dbl_kind) :: s, p
      real (kind=dbl_kind), dimension (n) :: work
      total = c0
      work(:) = c1
      do i = 1, n
         call inner (a(i), work(i), s, p)
         if (s > 3.0_dbl_kind) then
            a(i) = -a(i)**2
         elseif (s > 2.0_dbl_kind) then
            a(i) = s
         else
            a(i) = p
         endif
         if (a(i) < c0) total = total - a(i)
      enddo
      select case (trim(mode))
         case('double')
            total = total * 2.0_dbl_kind
         case('neg', 'minus')
            total = -total
         case default
            total = total
      end select
      end subroutine outer
""")


def test_subroutines_control_flow_and_out_arguments(tmp_path, monkeypatch):
    (tmp_path / "t.F90").write_text(FORTRAN)
    monkeypatch.setattr(rt, "REF", str(tmp_path))
    reg = {}
    reg["inner"] = rt.Sub("t.F90", "inner", reg)
    reg["outer"] = rt.Sub("t.F90", "outer", reg)
    assert reg["inner"].out_scalars == ["s", "p"] and reg["outer"].out_scalars == ["total"]
    assert reg["outer"].local_arrays == [("work", ["n"])]
    env = dict(ENV, c0=0.0, c1=1.0, _alloc=lambda n: rt.FArr(np.zeros(n)))
    for nm in ("inner", "outer"):
        exec(compile(reg[nm].python(), nm, "exec"), env)
    for mode, f in (("double", lambda t: t * 2.0), ("minus", lambda t: -t), ("other", lambda t: t)):
        a = np.array([0.5, 1.5, 2.5, 3.0])
        (total,) = env["outer"](4, rt.FArr(a), None, mode)
        # s = a+1: 1.5 -> p = 0.5; 2.5 -> s; 3.5 -> -(2.5**2); 4.0 -> -(9)
        assert a.tolist() == [0.5, 2.5, -6.25, -9.0]
        assert total == f(6.25 + 9.0)


def test_reference_constants_are_evaluated_from_the_source(tmp_path, monkeypatch):
    d = tmp_path / "cicecore" / "shared"
    d.mkdir(parents=True)
    (d / "ice_constants.F90").write_text(textwrap.dedent("""
      real (kind=dbl_kind), parameter, public :: &
        c1 = 1.0_dbl_kind, c9 = 9.0_dbl_kind, p5 = 0.5_dbl_kind, &
        p111 = c1/c9, &
        p055 = p111*p5
    """))
    monkeypatch.setattr(rt, "REF", str(tmp_path))
    k = rt.reference_constants()
    assert k["p111"] == 1.0 / 9.0 and k["p055"] == (1.0 / 9.0) * 0.5
