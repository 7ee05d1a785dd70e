#!/usr/bin/env python
"""Convert the reference's stored terminal ingredients into a plain numeric table.

Input  : /root/reference/ft_mpc/config/terminal.yaml  (read-only reference DATA file; its `cost`
         entry is a python/sympy expression string which the reference `eval`s at
         ft_mpc/controllers/tools/terminal_ingredients.py:451-474).
Output : fault-tolerant-mpc_b200/data/terminal.json

We never `eval` the string.  It is parsed with sympy.sympify on the inner expression and
decomposed into
    V_f(e) = sum_k  c_k * prod_i e_i^{p_ki}                       (polynomial terms)
           + sum_j  d_j * (prod_i e_i^{q_ji} + eps_j)^{w_j}       (smoothed-root terms)
with e = [ep1, ep2, ep3, ev1, ev2, ev3, eo1, eo2, eo3].
The decomposition is verified against direct sympy evaluation on random points, and the anchor
values are written into the JSON so that tests can pin every consumer (oracle and CUDA kernel).
"""
import json
import re
import sys
from pathlib import Path

import numpy as np
import sympy as sp
import yaml

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/ft_mpc/config/terminal.yaml")
OUT = Path(__file__).resolve().parent.parent / "fault-tolerant-mpc_b200" / "data" / "terminal.json"

NAMES = ["ep1", "ep2", "ep3", "ev1", "ev2", "ev3", "eo1", "eo2", "eo3"]
SYMS = [sp.Symbol(n) for n in NAMES]


def main():
    ing = yaml.safe_load(open(REF))
    src = " ".join(ing["cost"].split())
    # strip the "sp.lambdify((...), EXPR, modules={...})" wrapper textually
    m = re.match(r"^sp\.lambdify\(\((.*?)\),\s*(.*),\s*modules=\{.*\}\)\s*$", src)
    assert m, "unexpected layout of the cost string"
    assert [s.strip() for s in m.group(1).split(",")] == NAMES
    body = m.group(2)
    ns = {n: s for n, s in zip(NAMES, SYMS)}
    ns.update({"Float": sp.Float, "Symbol": sp.Symbol, "Abs": sp.Abs, "tanh": sp.tanh, "sqrt": sp.sqrt})
    expr = sp.sympify(body, locals=ns)
    expr = sp.expand(expr, power_base=False, power_exp=False, deep=False)

    poly_terms = {}
    root_terms = []
    const = 0.0

    def add_poly(coeff, exps):
        key = tuple(exps)
        poly_terms[key] = poly_terms.get(key, 0.0) + float(coeff)

    for term in sp.Add.make_args(expr):
        coeff, rest = term.as_coeff_Mul()
        factors = sp.Mul.make_args(rest)
        roots = [f for f in factors if isinstance(f, sp.Pow) and isinstance(f.base, sp.Add)]
        if roots:
            assert len(roots) == 1 and len(factors) == 1, f"unsupported term {term}"
            base, w = roots[0].base, float(roots[0].exp)
            eps, mono = base.as_coeff_Add()
            mc, mrest = mono.as_coeff_Mul()
            assert abs(float(mc) - 1.0) < 1e-15
            pd = sp.Poly(mrest, *SYMS).as_dict()
            assert len(pd) == 1
            (exps, c1), = pd.items()
            assert abs(float(c1) - 1.0) < 1e-15
            root_terms.append({"coeff": float(coeff), "exps": list(map(int, exps)), "eps": float(eps), "pow": w})
        else:
            pd = sp.Poly(term, *SYMS).as_dict()
            for exps, c in pd.items():
                if sum(exps) == 0:
                    const += float(c)
                else:
                    add_poly(c, exps)

    poly = [{"coeff": c, "exps": list(map(int, k))} for k, c in sorted(poly_terms.items()) if c != 0.0]

    tset = json.loads(ing["term_set"])
    A = np.array(tset["A"], dtype=float)
    b = np.array(tset["b"], dtype=float).reshape(-1)
    assert A.shape == (72, 9) and b.shape == (72,)

    # ---- verify the decomposition against sympy on random points
    f_ref = sp.lambdify(SYMS, expr, "mpmath")

    def f_tab(e):
        v = const
        for t in poly:
            v += t["coeff"] * np.prod(np.power(e, t["exps"]))
        for t in root_terms:
            v += t["coeff"] * (np.prod(np.power(e, t["exps"])) + t["eps"]) ** t["pow"]
        return v

    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(200):
        e = rng.uniform(-0.5, 0.5, 9)
        r = float(f_ref(*[sp.Float(x, 30) for x in e]))
        worst = max(worst, abs(r - f_tab(e)) / max(1.0, abs(r)))
    assert worst < 1e-12, worst

    anchors = []
    for e in [np.zeros(9), 0.1 * np.ones(9), np.array([.2, -.1, .05, -.3, .1, 0, .05, -.02, .1])]:
        anchors.append({"e": e.tolist(), "V": float(f_ref(*[sp.Float(x, 30) for x in e]))})

    out = {
        "_comment": "derived from the reference data file ft_mpc/config/terminal.yaml by tools/gen_terminal_data.py",
        "names": NAMES, "const": const, "poly": poly, "root": root_terms,
        "A": A.tolist(), "b": b.tolist(), "anchors": anchors,
    }
    OUT.write_text(json.dumps(out, indent=1))
    print(f"wrote {OUT}: {len(poly)} poly terms, {len(root_terms)} root terms, const {const!r}, "
          f"max rel err vs sympy {worst:.2e}")
    for a in anchors:
        print("anchor", a)


if __name__ == "__main__":
    main()
