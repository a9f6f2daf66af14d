"""Syntax side of the string macros (no GPU needed): Julia surface forms are translated the way @expression's
postwalk normalises them (lazyexpression.jl:141-160) and unhandled heads raise ArgumentError (:165-176)."""
import pytest

import parametron_b200 as pm
from parametron_b200 import macros


def test_translation_of_julia_surface_syntax():
    t = macros._split_dots(macros._translate("transpose(x) * Q * x + r ⋅ x + s"))
    assert t.replace(" ", "") == "transpose(x)*Q*x+dot(r,x)+s"
    assert macros._split_dots(macros._translate("resid' * resid")).replace(" ", "") == "resid.T*resid"
    assert macros._split_dots(macros._translate("[A * x; B * y]")).replace(" ", "") == "vcat(A*x,B*y)"
    assert macros._split_dots(macros._translate("w * (r ⋅ r) - p ⋅ x")).replace(" ", "") == "w*(dot(r,r))-dot(p,x)"


def test_unhandled_heads_raise_argument_error():
    class FakeModel:
        pass
    low = macros._Lower(FakeModel(), {"x": 1.0})
    for text in ("x if x else x", "lambda: x", "x[0]", "{x}", "x < 2", "'s'"):
        with pytest.raises(macros.ArgumentError):
            low.lower(text)
    with pytest.raises(macros.ArgumentError):
        macros.constraint(FakeModel(), "x + 1")            # no relation (model.jl:243)
    with pytest.raises(macros.ArgumentError):
        low.lower("y + 1")                                  # unknown name
    assert low.lower("2 * x + 1") == 3.0                    # constants fold on the host (lazyexpression.jl:189-192)
