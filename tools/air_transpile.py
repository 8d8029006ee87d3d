#!/usr/bin/env python3
"""Derives the symbolic AIR of a Cairo layout from the reference sources and serialises it as data.

    python tools/air_transpile.py [plain recursive starknet]     (build container only: reads /root/reference)

The reference defines each layout's constraints as Rust code that *builds* an expression DAG at run
time (`AirConfig::constraints(trace_len)`, layouts/src/<layout>/air.rs) from column enums whose
`offset()` impls encode the trace geometry.  This tool executes that construction symbolically — the
function body and the enum impls are mechanically rewritten to Python in memory and evaluated with a
symbolic trace length n — and writes the resulting DAG, independent of n, to
`sandstorm_b200/air/layouts/<layout>.json`:

    nodes: ["x"] | ["const", int] | ["gpow", a, d, c]        (= g^(a*n/d + c), g = generator of the trace domain)
         | ["trace", col, off] | ["challenge", i] | ["hint", i] | ["periodic", name, interval]
         | ["add"|"sub"|"mul"|"div", i, j] | ["neg", i] | ["pow", i, e] | ["xpow", a, d, c]   (= X^(a*n/d + c))
    constraints: node indices, in the order of the reference's constraint vector.

Nothing from the reference is copied into the repository as source: the committed artefact is the DAG
(generated data, like tests/golden/*), and this script is how it is regenerated.
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = os.environ.get("SANDSTORM_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "sandstorm_b200", "air", "layouts")
P = 2**251 + 17 * 2**192 + 1


# ------------------------------------------------------------------------------ symbolic integers
class Sym:
    """a*n/d + c  (exact); n = trace length."""

    def __init__(self, a=0, d=1, c=0):
        from math import gcd
        g = gcd(a, d) or 1
        self.a, self.d, self.c = a // g, d // g, c

    def _lift(o):
        return o if isinstance(o, Sym) else Sym(0, 1, int(o))

    def __add__(self, o):
        o = Sym._lift(o)
        return Sym(self.a * o.d + o.a * self.d, self.d * o.d, self.c + o.c)
    __radd__ = __add__

    def __sub__(self, o):
        o = Sym._lift(o)
        return Sym(self.a * o.d - o.a * self.d, self.d * o.d, self.c - o.c)

    def __rsub__(self, o):
        return Sym._lift(o) - self

    def __mul__(self, o):
        if isinstance(o, Sym):
            if o.a and self.a:
                raise ValueError("n*n")
            k, s = (o.c, self) if not o.a else (self.c, o)
            return Sym(s.a * k, s.d, s.c * k)
        return Sym(self.a * int(o), self.d, self.c * int(o))
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = int(o) if not isinstance(o, Sym) else (o.c if not o.a else None)
        if o is None or self.c % o:
            raise ValueError("inexact symbolic division")
        return Sym(self.a, self.d * o, self.c // o)
    __floordiv__ = __truediv__

    def __rtruediv__(self, o):
        if self.a:
            raise ValueError("division by symbolic n")
        return I(int(o) // self.c)

    def key(self):
        return (self.a, self.d, self.c)

    def __repr__(self):
        return f"Sym({self.a}n/{self.d}+{self.c})"


class I(int):
    """usize semantics: '/' is integer division."""

    def __truediv__(self, o):
        if isinstance(o, Sym):
            return o.__rtruediv__(self)
        return I(int(self) // int(o))

    def __rtruediv__(self, o):
        return I(int(o) // int(self))

    def __mul__(self, o):
        return o.__rmul__(self) if isinstance(o, Sym) else I(int(self) * int(o))
    __rmul__ = __mul__

    def __add__(self, o):
        return o.__radd__(self) if isinstance(o, Sym) else I(int(self) + int(o))
    __radd__ = __add__

    def __sub__(self, o):
        return o.__rsub__(self) if isinstance(o, Sym) else I(int(self) - int(o))

    def __rsub__(self, o):
        return I(int(o) - int(self))

    def __neg__(self):
        return I(-int(self))

    def pow(self, e):
        return I(int(self) ** int(e))


# ------------------------------------------------------------------------------ expression nodes
class Node:
    table: dict = {}
    nodes: list = []

    def __new__(cls, *key):
        hit = cls.table.get(key)
        if hit is not None:
            return hit
        self = object.__new__(cls)
        self.key, self.idx = key, len(cls.nodes)
        cls.table[key] = self
        cls.nodes.append(self)
        return self

    @classmethod
    def reset(cls):
        cls.table, cls.nodes = {}, []

    @staticmethod
    def lift(v):
        if isinstance(v, Node):
            return v
        if isinstance(v, GPow):
            return Node("gpow", *v.e.key())
        if isinstance(v, Sym):
            raise ValueError("symbolic integer used as a field constant")
        return Node("const", int(v) % P)

    def _bin(self, op, o, swap=False):
        o = Node.lift(o)
        a, b = (o, self) if swap else (self, o)
        return Node(op, a.idx, b.idx)

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __neg__(self): return Node("neg", self.idx)

    def pow(self, e):
        if self.key == ("x",):
            e = e if isinstance(e, Sym) else Sym(0, 1, int(e))
            return Node("xpow", *e.key())
        if isinstance(e, Sym):
            if e.a:
                raise ValueError("symbolic exponent on a non-X base")
            e = e.c
        return Node("pow", self.idx, int(e))

    def clone(self): return self
    def square(self): return self * self


class GPow:
    """g^e, g = trace-domain generator, e symbolic."""

    def __init__(self, e):
        self.e = e if isinstance(e, Sym) else Sym(0, 1, int(e))

    def pow(self, k):
        return GPow(self.e * (k if isinstance(k, Sym) else int(k)))

    def __mul__(self, o):
        if isinstance(o, GPow):
            return GPow(self.e + o.e)
        return Node.lift(self) * o

    def inverse(self):
        return _Opt(GPow(Sym(0, 1, 0) - self.e))


class _Opt:
    def __init__(self, v): self.v = v
    def unwrap(self): return self.v


class _Domain:
    def group_gen(self): return GPow(1)


# ------------------------------------------------------------------------------ Rust text helpers
def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def match_brace(src: str, open_pos: int, o="{", c="}") -> int:
    depth = 0
    for i in range(open_pos, len(src)):
        if src[i] == o:
            depth += 1
        elif src[i] == c:
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced")


def split_top(src: str, sep: str) -> list[str]:
    out, depth, cur = [], 0, []
    for ch in src:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur)); cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur))
    return out


def rewrite_match(expr: str, self_name: str) -> str:
    """`match self { A | B => e1, C => { e2 } }`  ->  nested Python conditional."""
    while True:
        m = re.search(r"\bmatch\s+(\*?self)\s*\{", expr)
        if not m:
            return expr
        close = match_brace(expr, m.end() - 1)
        arms = []
        inner = expr[m.end():close]
        pos = 0
        while True:
            arrow = inner.find("=>", pos)
            if arrow < 0:
                break
            pats = inner[pos:arrow]
            k = arrow + 2
            while k < len(inner) and inner[k].isspace():
                k += 1
            if k < len(inner) and inner[k] == "{":          # block arm: the comma after it is optional
                end = match_brace(inner, k)
                body = inner[k + 1:end].strip()
                pos = end + 1
                while pos < len(inner) and (inner[pos].isspace() or inner[pos] == ","):
                    pos += 1
            else:
                depth, j = 0, k
                while j < len(inner) and not (inner[j] == "," and depth == 0):
                    depth += inner[j] in "([{"
                    depth -= inner[j] in ")]}"
                    j += 1
                body = inner[k:j].strip()
                pos = j + 1
            names = [p.strip().split("::")[-1] for p in pats.split("|")]
            arms.append((names, rewrite_match(body, self_name)))
        py = "_nomatch()"
        for names, body in reversed(arms):
            cond = "True" if names == ["_"] else " or ".join(f"{self_name}.name == '{n}'" for n in names)
            py = f"(({body}) if ({cond}) else {py})"
        expr = expr[:m.start()] + py + expr[close + 1:]


def rewrite_expr(e: str) -> str:
    e = e.replace("\n", " ")
    e = re.sub(r"\b(unimplemented|unreachable|todo|panic)!\([^)]*\)", "_nomatch()", e)
    e = re.sub(r"\b([A-Z]\w*::[A-Z]\w*)\s+as\s+(u64|usize|isize|u32)\b", r"\1.value", e)     # enum discriminant casts
    e = re.sub(r"\bas\s+(u64|usize|isize|u32|i64|u128|u8)\b", "", e)
    e = re.sub(r"(\d+)(u8|u16|u32|u64|usize|isize|i32|i64)\b", r"I(\1)", e)
    while True:                                               # x.pow([e])  ->  x.pow((e))
        pm = re.search(r"\.pow\(\s*\[", e)
        if not pm:
            break
        close = match_brace(e, pm.end() - 1, "[", "]")
        e = e[:pm.start()] + ".pow((" + e[pm.end():close] + ")" + e[close + 1:]
    e = re.sub(r"BigUint::from\(", "_id(", e)
    e = e.replace("Expr::from(", "_id(").replace("FieldVariant::Fp(", "_id(").replace("Fp::from(", "_id(")
    e = e.replace("Fp::ONE", "I(1)").replace("Fp::ZERO", "I(0)").replace("Fp::one()", "I(1)").replace("Fp::zero()", "I(0)")
    e = re.sub(r'MontFp!\(\s*"(\d+)"\s*\)', r"I(\1)", e)
    e = re.sub(r"::<[A-Za-z0-9_, ]+>", "", e)                  # turbofish
    e = re.sub(r"Radix2EvaluationDomain::new\([^)]*\)\.unwrap\(\)", "_Domain()", e)
    e = re.sub(r"&\s*(?=[A-Za-z_(])", "", e)
    e = re.sub(r"\*self\b", "self.value", e)
    e = re.sub(r"\b([A-Z][A-Za-z0-9_]*)::([A-Za-z_][A-Za-z0-9_]*)", r"\1.\2", e)
    e = re.sub(r"\b(pedersen|ecdsa|poseidon)::", r"\1.", e)
    e = re.sub(r"\b(constants|params)::", r"\1.", e)
    e = re.sub(r"\bsuper::", "", e)
    e = re.sub(r"\)\.([01])\b", r")[\1]", e)                   # tuple field access
    return e


# ------------------------------------------------------------------------------ enums
def make_enums(src: str, env: dict):
    src_nc = strip_comments(src)
    enums = {}
    for m in re.finditer(r"pub enum (\w+)\s*\{", src_nc):
        name = m.group(1)
        body = src_nc[m.end():match_brace(src_nc, m.end() - 1)]
        variants, nxt = [], 0
        for item in split_top(body, ","):
            item = item.strip()
            if not item:
                continue
            item = re.sub(r"#\[[^\]]*\]", "", item).strip()
            if "=" in item:
                v, d = item.split("=")
                nxt = int(eval(rewrite_expr(d), dict(env)))
                variants.append((v.strip(), nxt))
            else:
                variants.append((item, nxt))
            nxt += 1
        enums[name] = variants

    methods: dict[str, dict[str, str]] = {n: {} for n in enums}
    for m in re.finditer(r"impl(?:\s+(\w+)(?:<[^>]*>)?\s+for)?\s+(\w+)\s*\{", src_nc):
        trait, name = m.group(1), m.group(2)
        if name not in enums:
            continue
        body = src_nc[m.end():match_brace(src_nc, m.end() - 1)]
        for f in re.finditer(r"fn (\w+)(?:<[^>]*>)?\s*\(([^)]*)\)[^{]*\{", body):
            fname, params = f.group(1), f.group(2)
            fbody = body[f.end():match_brace(body, f.end() - 1)]
            args = [p.split(":")[0].strip() for p in params.split(",") if p.strip() and "self" not in p.split(":")[0]]
            stmts = [s.strip() for s in split_top(fbody, ";") if s.strip()]
            lines = []
            for k, s in enumerate(stmts):
                if s.startswith("use "):
                    continue
                s = rewrite_expr(rewrite_match(s, "self"))
                s = s.replace("AlgebraicItem.Trace(", "Trace(").replace(".into()", "")
                lm = re.match(r"let\s+(\(?[\w\s,]+\)?)\s*(?::[^=]+)?=\s*(.*)$", s, flags=re.S)
                if lm:
                    lines.append(f"    {lm.group(1).strip()} = {lm.group(2)}")
                elif k == len(stmts) - 1:
                    lines.append(f"    return {s}")
                else:
                    lines.append(f"    {s}")
            code = f"def {fname}(self{''.join(', ' + a for a in args)}):\n" + "\n".join(lines) + "\n"
            methods[name][fname] = code
            methods[name].setdefault("_traits", "")
            methods[name]["_traits"] += f" {trait}"

    classes = {}
    for name, variants in enums.items():
        ns = dict(env)
        cls = type(name, (), {})
        for fname, code in methods[name].items():
            if fname == "_traits":
                continue
            loc = {}
            exec(code, ns, loc)
            setattr(cls, fname, loc[fname])
        traits = methods[name].get("_traits", "")
        if "ExecutionTraceColumn" in traits:
            cls.curr = lambda self: self.offset(0)
            cls.next = lambda self: self.offset(1)
        if "VerifierChallenge" in traits:
            cls.challenge = lambda self: Node("challenge", int(self.index()))
        if "Hint" in traits:
            cls.hint = lambda self: Node("hint", int(self.index()))
        for v, d in variants:
            inst = cls()
            inst.name, inst.value = v, I(d)
            setattr(cls, v, inst)
        classes[name] = cls
        ns[name] = cls
        env[name] = cls
    for cls in classes.values():          # methods may reference sibling enums
        for attr in list(vars(cls).values()):
            if callable(attr) and hasattr(attr, "__globals__"):
                attr.__globals__.update(classes)
    return classes


# ------------------------------------------------------------------------------ main transpile
def load_consts(layout: str) -> dict:
    src = strip_comments(open(os.path.join(REF, "layouts/src", layout, "mod.rs")).read())
    env = {"I": I}
    for m in re.finditer(r"pub const (\w+): usize = ([^;]+);", src):
        env[m.group(1)] = I(eval(rewrite_expr(m.group(2)), dict(env)))
    return env


def periodic_columns(src: str, env: dict) -> dict:
    out = {}
    for m in re.finditer(r"const (\w+): PeriodicColumn<[^=]*=\s*\{", src):
        body = src[m.end():match_brace(src, m.end() - 1)]
        interval = re.search(r"INTERVAL_SIZE: usize = ([^;]+);", body).group(1)
        coeffs = re.search(r"map_into_fp_array\(\s*\w+::periodic::(\w+)\s*\)", body).group(1)
        out[m.group(1)] = (coeffs, int(eval(rewrite_expr(interval), dict(env))))
    return out


class _NS:
    def __init__(self, **kw): self.__dict__.update(kw)


def builtin_constants() -> dict:
    """Numeric constants of the builtins crate that the AIRs reference by path."""
    ped = strip_comments(open(os.path.join(REF, "builtins/src/pedersen/constants.rs")).read())
    pts = {}
    for m in re.finditer(r'pub const (P\d): Affine<StarkwareCurve> = Affine::new_unchecked\(\s*Fp!\("(\d+)"\),\s*Fp!\("(\d+)"\),?\s*\)', ped):
        pts[m.group(1)] = _NS(x=I(int(m.group(2))), y=I(int(m.group(3))))
    pos = strip_comments(open(os.path.join(REF, "builtins/src/poseidon/params.rs")).read())

    def table(name):
        m = re.search(r"pub const %s: \[\[Fp; 3\]; [^\]]+\] = \[(.*?)\n\];" % name, pos, re.S)
        vals = [I(int(v)) for v in re.findall(r'Fp!\(\s*"(\d+)"\s*\)', m.group(1))]
        return [vals[3 * i:3 * i + 3] for i in range(len(vals) // 3)]

    first, partial, second = table("FULL_ROUND_KEYS_1ST_HALF"), table("PARTIAL_ROUND_KEYS"), table("FULL_ROUND_KEYS_2ND_HALF")
    curve = strip_comments(open(os.path.join(REF, "builtins/src/utils.rs")).read())
    beta = I(int(re.search(r'COEFF_B: Self::BaseField =\s*Fp!\("(\d+)"\)', curve).group(1)))
    return {
        "pedersen": _NS(constants=_NS(**pts)),
        "ecdsa": _NS(SHIFT_POINT=pts["P0"]),
        "poseidon": _NS(params=_NS(ROUND_KEYS=first + partial + second, PARTIAL_ROUND_KEYS=partial,
                                   FULL_ROUND_KEYS_1ST_HALF=first, FULL_ROUND_KEYS_2ND_HALF=second)),
        "ECDSA_SIG_CONFIG_ALPHA": I(1), "ECDSA_SIG_CONFIG_BETA": beta,
    }


def _nomatch():
    raise ValueError("non-exhaustive match")


def transpile(layout: str) -> dict:
    Node.reset()
    src = open(os.path.join(REF, "layouts/src", layout, "air.rs")).read()
    env = load_consts(layout)
    periodic = periodic_columns(strip_comments(src), env)
    env.update({
        "Trace": lambda col, off: Node("trace", int(col), int(off)),
        "X": Node("x"), "_Domain": _Domain, "_nomatch": _nomatch, "Sym": Sym,
        "_id": lambda v: Node.lift(v) if isinstance(v, (Node, GPow)) else v,
        "Constant": lambda v: Node.lift(v),
        "Periodic": lambda spec: Node("periodic", spec[0], spec[1]),
    })
    env.update(periodic)
    env.update(builtin_constants())
    make_enums(src, env)

    nc = strip_comments(src)
    m = re.search(r"fn constraints\(trace_len: usize\)[^{]*\{", nc)
    body = nc[m.end():match_brace(nc, m.end() - 1)]
    env["trace_len"] = Sym(1, 1, 0)
    for um in re.finditer(r"use (\w+)::\*;", body):
        cls = env.get(um.group(1))
        if isinstance(cls, type):
            for k, v in vars(cls).items():
                if hasattr(v, "name") and hasattr(v, "value"):
                    env[k] = v
    result = None
    for stmt in split_top(body, ";"):
        s = stmt.strip()
        if not s or s.startswith("use ") or s.startswith("assert"):
            continue
        s = rewrite_expr(s)
        lm = re.match(r"let\s+(mut\s+)?(\w+)\s*(?::[^=]+)?=\s*(.*)$", s, flags=re.S)
        if lm:
            try:
                env[lm.group(2)] = eval(lm.group(3), env)
            except Exception as exc:
                raise RuntimeError(f"while evaluating `let {lm.group(2)}` = {lm.group(3)[:300]}") from exc
            continue
        if s.startswith("vec!["):
            inner = s[s.index("[") + 1:match_brace(s, s.index("["), "[", "]")]
            result = [eval(x.strip(), env) for x in split_top(inner, ",") if x.strip()]
            continue
        raise ValueError(f"unhandled statement: {s[:120]}")
    nodes = [list(nd.key) for nd in Node.nodes]
    cons = [Node.lift(c).idx for c in result]
    # keep only nodes reachable from the constraints
    keep, stack = set(), list(cons)
    while stack:
        i = stack.pop()
        if i in keep:
            continue
        keep.add(i)
        k = nodes[i]
        if k[0] in ("add", "sub", "mul", "div"):
            stack += [k[1], k[2]]
        elif k[0] in ("neg", "pow"):
            stack.append(k[1])
    order = sorted(keep)
    remap = {old: new for new, old in enumerate(order)}
    packed = []
    for old in order:
        k = list(nodes[old])
        if k[0] in ("add", "sub", "mul", "div"):
            k[1], k[2] = remap[k[1]], remap[k[2]]
        elif k[0] in ("neg", "pow"):
            k[1] = remap[k[1]]
        elif k[0] == "const":
            k[1] = hex(k[1])
        packed.append(k)
    taps = sorted({(k[1], k[2]) for k in packed if k[0] == "trace"})
    # the AirConfig impl is authoritative for the column counts (recursive/mod.rs:29-30 is stale: SURVEY App. B)
    for cname in ("NUM_BASE_COLUMNS", "NUM_EXTENSION_COLUMNS"):
        cm = re.search(r"const %s: usize = (\d+);" % cname, nc)
        if cm:
            env[cname] = I(int(cm.group(1)))
    return {
        "layout": layout,
        "source": f"generated by tools/air_transpile.py from layouts/src/{layout}/air.rs (AirConfig::constraints)",
        "num_base_columns": int(env["NUM_BASE_COLUMNS"]), "num_extension_columns": int(env["NUM_EXTENSION_COLUMNS"]),
        "cycle_height": int(env["CYCLE_HEIGHT"]),
        "periodic": {name: {"coeffs": spec[0], "interval": spec[1]} for name, spec in periodic.items()},
        "n_constraints": len(cons), "n_taps": len(taps), "max_offset": max(t[1] for t in taps),
        "nodes": packed, "constraints": [remap[c] for c in cons],
    }


def main():
    layouts = sys.argv[1:] or ["plain", "recursive", "starknet"]
    os.makedirs(OUT_DIR, exist_ok=True)
    for layout in layouts:
        data = transpile(layout)
        with open(os.path.join(OUT_DIR, f"{layout}.json"), "w") as f:
            json.dump(data, f, separators=(",", ":"))
        per_col = {}
        for k in data["nodes"]:
            if k[0] == "trace":
                per_col[k[1]] = per_col.get(k[1], 0) + 1
        print(layout, "constraints", data["n_constraints"], "nodes", len(data["nodes"]), "taps", data["n_taps"],
              "max_offset", data["max_offset"], "taps/col", [per_col.get(c, 0) for c in range(max(per_col) + 1)])
    # periodic coefficient tables used by the layouts (the same file tests/golden/make_golden.py extracts)
    import shutil
    shutil.copy(os.path.join(ROOT, "tests", "golden", "periodic_coeffs.json"), os.path.join(OUT_DIR, "periodic_coeffs.json"))


if __name__ == "__main__":
    main()
