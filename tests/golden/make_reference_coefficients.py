"""Generates tests/golden/reference_coefficients.json: the coefficient tables of bacon_sci::ivp READ OUT OF THE REFERENCE'S
OWN SOURCE TEXT (/root/reference/src/ivp/{rk,bdf,adams}.rs), so that the oracle's restatement of them is pinned on the
reference itself and not on a second reading of it.  The reference is Rust and cannot be executed here; its tables are
literal expressions over `RealField::from_u8/u16/f64`, `zero()`, `one()`, `.recip()`, `/` and unary minus, which this
script evaluates in IEEE double exactly as f64 would (a / b, 1.0 / x).  Lists are kept IN SOURCE ORDER: what
`BSMatrix::from_vec` makes of a list (column-major fill, SURVEY.md defect D1) is the oracle's business, not this file's.
Run from the repo root in the build container (the GPU box has no /root/reference):
    python tests/golden/make_reference_coefficients.py
"""
import json
import os
import re

REF = "/root/reference/src/ivp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_coefficients.json")


def function_body(text, impl_pattern, fn_name):
    """(body, first line number) of `fn fn_name` inside the first impl block matching impl_pattern."""
    m = re.search(impl_pattern, text)
    assert m, impl_pattern
    f = re.compile(r"fn\s+" + fn_name + r"\s*\(").search(text, m.end())
    assert f, fn_name
    i = text.index("{", f.end())
    depth, j = 0, i
    while True:
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return text[i + 1:j], text.count("\n", 0, i) + 1


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def split_top_level(s):
    """split on commas that are not inside parentheses"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out if x.strip()]


def evaluate(expr, env):
    """One table entry: [-] term [/ term], term = from_uN(k)? | from_f64(<float expr>).unwrap() | zero() | one() | name,
    each optionally followed by .recip() and / or .clone().  Method calls bind tighter than unary minus, as in Rust."""
    e = expr.strip()
    depth = 0
    for i, ch in enumerate(e):  # a top-level sum: a + b (adams.rs:642)
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "+" and depth == 0 and i > 0:
            return evaluate(e[:i], env) + evaluate(e[i + 1:], env)
    neg = e.startswith("-")
    if neg:
        e = e[1:].strip()
    depth, cut = 0, None
    for i, ch in enumerate(e):  # the top-level division, if any
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "/" and depth == 0:
            cut = i
            break
    if cut is not None:
        v = term(e[:cut], env) / term(e[cut + 1:], env)
    else:
        v = term(e, env)
    return -v if neg else v


def term(t, env):
    t = t.strip().replace(".clone()", "")
    recip = t.endswith(".recip()")
    if recip:
        t = t[:-len(".recip()")]
    t = t.strip()
    m = re.fullmatch(r"(?:Self::RealField|N::RealField|Self::Field)::from_u(?:8|16|32)\((\d+)\)(?:\.ok_or\(IVPError::FromPrimitiveFailure\))?\?", t)
    if m:
        v = float(int(m.group(1)))
    elif re.fullmatch(r"(?:Self::RealField|N::RealField)::zero\(\)", t):
        v = 0.0
    elif re.fullmatch(r"(?:Self::RealField|N::RealField)::one\(\)", t):
        v = 1.0
    else:
        m = re.fullmatch(r"(?:Self::RealField|N::RealField)::from_f64\((.*)\)\.unwrap\(\)", t)
        if m:
            assert re.fullmatch(r"[-0-9. /]+", m.group(1)), m.group(1)
            v = float(eval(m.group(1)))  # a literal f64 expression such as -128.0 / 4275.0
        else:
            assert t in env, f"unknown term {t!r}"
            v = env[t]
    return 1.0 / v if recip else v


def table(path, impl_pattern, fn_name):
    text = open(path).read()
    body, line = function_body(text, impl_pattern, fn_name)
    body = strip_comments(body)
    env = {}
    for name, rhs in re.findall(r"let\s+(\w+)\s*=\s*([^;]+);", body):
        env[name] = evaluate(rhs, env)
    m = re.search(r"(?:from_column_slice\(&\[|from_vec\(vec!\[)(.*?)\]\)", body, re.S)
    if m:
        values = [evaluate(x, env) for x in split_top_level(m.group(1))]
    else:  # a single value: Some(expr)
        m = re.search(r"Some\((.*)\)\s*$", body.strip(), re.S)
        values = [evaluate(m.group(1), env)]
    return {"source": f"{os.path.basename(path)}:{line}", "values": values}


def scalar_binding(path, name, needs):
    """value of `let name = <expr>;` where it first appears in the file, with the `needs` bindings it is built from
    (the step-size safety factor: rk.rs:266-268 builds "eighty_four" from 100 — SURVEY.md defect D3)"""
    text = strip_comments(open(path).read())
    env, texts = {}, {}
    for nm in list(needs) + [name]:
        m = re.search(r"let\s+" + nm + r"\s*=\s*([^;]+);", text)
        assert m, nm
        env[nm] = evaluate(m.group(1), env)
        texts[nm] = " ".join(m.group(1).split())
        line = text.count(chr(10), 0, m.start()) + 1
    return {"source": f"{os.path.basename(path)}:{line}", "as_written": texts, "values": [env[name]]}


def main():
    rk, bdf, adams = (os.path.join(REF, f) for f in ("rk.rs", "bdf.rs", "adams.rs"))
    out = {"_generated_by": "tests/golden/make_reference_coefficients.py from /root/reference/src/ivp (lists in SOURCE order)"}
    for key, pat in (("RK45", r"impl<N: ComplexField> RungeKuttaCoefficients<6> for"), ("RK23", r"impl<N: ComplexField> RungeKuttaCoefficients<4> for")):
        out[key] = {f: table(rk, pat, f) for f in ("t_coefficients", "k_coefficients", "avg_coefficients", "error_coefficients")}
    out["RK_safety"] = scalar_binding(rk, "point_eighty_four", ("one_hundred", "eighty_four"))
    for key, pat in (("BDF6", r"BDFCoefficients<7> for"), ("BDF2", r"BDFCoefficients<3> for")):
        out[key] = {f: table(bdf, pat, f) for f in ("higher_coefficients", "lower_coefficients")}
    for key, pat in (("Adams5", r"AdamsCoefficients<5> for"), ("Adams3", r"AdamsCoefficients<3> for")):
        out[key] = {f: table(adams, pat, f) for f in ("predictor_coefficients", "corrector_coefficients", "error_coefficient")}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        if isinstance(v, dict) and "values" not in v:
            print(k, {f: (t["source"], len(t["values"])) for f, t in v.items()})
        else:
            print(k, v)


if __name__ == "__main__":
    main()
