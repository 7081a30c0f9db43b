"""Parser for the Maple-generated straight-line programs (SLPs) of the MB-pol PIPs.

Reads the reference's generated sources *where they lie* (never copied into this
repo) and turns them into a DAG that our own generators re-schedule and re-emit:

  /root/reference/platforms/reference/src/poly-2b-v6x.cpp:3-13841   (E, g[31]; 1153 coeffs)
  /root/reference/platforms/reference/src/poly-3b-v2x.cpp:3-28378   (E, g[36]; 1163 coeffs)

Node kinds (integer ids index the parallel lists of ``Dag``):
  'a'  coefficient  (arg = index)       'x'  variable (arg = index)
  'c'  numeric constant (arg = value)   'add' / 'mul'  n-ary, args = child ids in source order

Outputs: ``dag.outputs['E']`` and ``dag.outputs['g'][i]``.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

REF_SRC = "/root/reference/platforms/reference/src"
POLY2B = f"{REF_SRC}/poly-2b-v6x.cpp"
POLY3B = f"{REF_SRC}/poly-3b-v2x.cpp"

_TOK = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?|\d+)|([A-Za-z_]\w*)|(.))")


@dataclass
class Dag:
    kind: list = field(default_factory=list)   # str
    args: list = field(default_factory=list)   # tuple[int,...] | int | float
    outputs: dict = field(default_factory=dict)
    n_a: int = 0
    n_x: int = 0
    _leaf: dict = field(default_factory=dict)

    def leaf(self, kind, arg):
        key = (kind, arg)
        nid = self._leaf.get(key)
        if nid is None:
            nid = self.new(kind, arg)
            self._leaf[key] = nid
        return nid

    def new(self, kind, args):
        self.kind.append(kind)
        self.args.append(args)
        return len(self.kind) - 1

    def __len__(self):
        return len(self.kind)

    # ---- statistics -------------------------------------------------------------------------
    def live_nodes(self):
        """ids reachable from the outputs, in topological (id) order."""
        seen = set()
        stack = [self.outputs["E"], *self.outputs["g"]]
        while stack:
            n = stack.pop()
            if n in seen:
                continue
            seen.add(n)
            if self.kind[n] in ("add", "mul"):
                stack.extend(self.args[n])
        return sorted(seen)

    def op_counts(self):
        mul = add = 0
        for n in self.live_nodes():
            k = self.kind[n]
            if k == "add":
                add += len(self.args[n]) - 1
            elif k == "mul":
                mul += len(self.args[n]) - 1
        return mul, add


class _Parser:
    def __init__(self, text, dag, env):
        self.toks = []
        for m in _TOK.finditer(text):
            num, ident, sym = m.groups()
            if num is not None:
                self.toks.append(("num", float(num)))
            elif ident is not None:
                self.toks.append(("id", ident))
            elif sym is not None and not sym.isspace():
                self.toks.append(("sym", sym))
        self.i = 0
        self.dag = dag
        self.env = env

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else ("eof", None)

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def expect(self, sym):
        t = self.next()
        if t != ("sym", sym):
            raise SyntaxError(f"expected {sym!r}, got {t!r}")

    def expr(self):
        terms = [self.term()]
        while self.peek() == ("sym", "+"):
            self.next()
            terms.append(self.term())
        if len(terms) == 1:
            return terms[0]
        return self.dag.new("add", tuple(terms))

    def term(self):
        facs = [self.factor()]
        while self.peek() == ("sym", "*"):
            self.next()
            facs.append(self.factor())
        if len(facs) == 1:
            return facs[0]
        return self.dag.new("mul", tuple(facs))

    def factor(self):
        kind, val = self.next()
        if kind == "num":
            return self.dag.leaf("c", val)
        if kind == "sym" and val == "(":
            e = self.expr()
            self.expect(")")
            return e
        if kind == "id":
            if self.peek() == ("sym", "["):
                self.next()
                k, idx = self.next()
                assert k == "num"
                idx = int(idx)
                self.expect("]")
                if val == "a":
                    self.dag.n_a = max(self.dag.n_a, idx + 1)
                    return self.dag.leaf("a", idx)
                if val == "x":
                    self.dag.n_x = max(self.dag.n_x, idx + 1)
                    return self.dag.leaf("x", idx)
                return self.env[(val, idx)]          # df[k]: most recent assignment
            return self.env[val]
        raise SyntaxError(f"unexpected token {(kind, val)!r}")


def parse(path):
    """Parse one generated polynomial source into a Dag."""
    src = open(path).read()
    body = src[src.index("{") + 1: src.rindex("}")]
    dag = Dag()
    env = {}
    g = {}
    for stmt in body.split(";"):
        s = stmt.strip()
        if not s:
            continue
        if s.startswith("double df["):          # declaration of the adjoint scratch array
            continue
        if s.startswith("return"):
            dag.outputs["E"] = _Parser(s[len("return"):], dag, env).expr()
            continue
        lhs, rhs = s.split("=", 1)
        lhs = lhs.replace("const double", "").strip()
        node = _Parser(rhs, dag, env).expr()
        m = re.fullmatch(r"(\w+)\[(\d+)\]", lhs)
        if m:
            name, idx = m.group(1), int(m.group(2))
            if name == "g":
                g[idx] = node
            else:
                env[(name, idx)] = node
        else:
            env[lhs] = node
    dag.outputs["g"] = [g[i] for i in range(len(g))]
    return dag


if __name__ == "__main__":
    import sys
    for p in (POLY2B, POLY3B):
        d = parse(p)
        live = d.live_nodes()
        print(p.split("/")[-1], "nodes", len(d), "live", len(live), "mul/add", d.op_counts(),
              "n_a", d.n_a, "n_x", d.n_x, file=sys.stderr)
