#!/usr/bin/env python
"""v2c.py -- TEST INFRASTRUCTURE ONLY: Verilog-subset -> C translator for the reference's own source.

The reference (linfenghuaster/Regex-FPGA) is Verilog and this image has no HDL simulator.  This tool makes
the reference's UNMODIFIED source executable anyway: it parses Design/FPGA.v where it lies (the file is
never copied into this repository) and emits a C model of the module with the language's own simulation
semantics, which oracle/vsim/tb_driver.c clocks exactly as Simulation/testbench_BLK_Mem.sv does.  The
result (oracle/_ref/libref.so, built by `make -C oracle _ref`) pins the hand-written oracles: tests/
compare it cycle by cycle with oracle A.

Subset (everything Design/FPGA.v uses; anything else is a hard error, never silently skipped):
  * one module; `input` / `output` / `reg` / `integer` / `parameter` declarations; memories (`reg [..] m [..]`)
  * `always @(posedge clk)` with non-blocking assignments, `always @(*)` with blocking assignments
  * begin/end, if / else if / else
  * lvalues: r, r[expr], r[c:c], m[expr]; expressions: numbers incl. sized/based literals with x/z digits,
    names, bit- and part-selects (also of memory words: m[k][31:24]), ! ~ unary -, + - << >> < <= > >= == !=
    && ||, parentheses
Semantics implemented:
  * non-blocking: every right-hand side of an edge reads pre-edge values; updates are committed after the
    whole block was evaluated, in program order (last write wins); bit / part / whole-vector updates of
    wide vectors go through an ordered update queue whose sources are snapshotted before any is applied
  * blocking assignments in `always @(*)` blocks are evaluated in source order before every clock edge
    (their inputs only change at clock edges); a variable a path does not assign keeps its value (latch)
  * values are 2-state.  x/z digits of literals and the power-up value of every reg are filled with a
    caller-chosen bit (`xfill` = 0 or 1): tests run both fills and require identical outputs, which shows
    that no x/z ever reaches anything the testbench observes (IEEE 1364 would propagate X there)
  * arithmetic is carried out in 64 bits and truncated to the width of the assigned reg; that equals
    Verilog's context-determined sizing for every expression form of the subset as long as no
    intermediate exceeds 64 bits (operands are at most 25 bits wide in FPGA.v)
  * an out-of-range bit-select reads 0 and an out-of-range bit write is ignored (IEEE 1364: x / no effect)
Parameters named with --runtime-param become fields set at construction (size_range: the testbench
overrides it per ruleset, testbench_BLK_Mem.sv:20,94); all others are folded at translation time.
"""
import argparse
import re
import sys

# ------------------------------------------------------------------------------------------------
# lexer
# ------------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lcomment>//[^\n]*)
  | (?P<bcomment>/\*.*?\*/)
  | (?P<directive>`[a-zA-Z_]+[^\n]*)
  | (?P<based>(?:\d+\s*)?'[sS]?[bBdDhHoO]\s*[0-9a-fA-FxXzZ_?]+)
  | (?P<num>\d[\d_]*)
  | (?P<id>[A-Za-z_][A-Za-z0-9_$]*)
  | (?P<op><=|>=|==|!=|&&|\|\||<<|>>|[-+*/%<>!~&|^?:;,.()\[\]{}=@\#])
""", re.X | re.S)

KEYWORDS = {"module", "endmodule", "input", "output", "reg", "integer", "parameter", "always", "posedge",
            "begin", "end", "if", "else", "wire"}


class Tok:
    def __init__(self, kind, text, line):
        self.kind, self.text, self.line = kind, text, line

    def __repr__(self):
        return f"{self.kind}:{self.text}@{self.line}"


def lex(src):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"line {line}: cannot tokenise {src[pos:pos + 20]!r}")
        kind, text = m.lastgroup, m.group()
        if kind in ("num", "based", "id", "op"):
            if kind == "id" and text in KEYWORDS:
                kind = "kw"
            toks.append(Tok(kind, text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(Tok("eof", "", line))
    return toks


# ------------------------------------------------------------------------------------------------
# parser -> AST (tuples)
# ------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, toks):
        self.t, self.p = toks, 0

    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def accept(self, text):
        if self.peek().text == text and self.peek().kind in ("op", "kw"):
            return self.next()
        return None

    def expect(self, text):
        tok = self.next()
        if tok.text != text:
            raise SyntaxError(f"line {tok.line}: expected {text!r}, found {tok.text!r}")
        return tok

    def ident(self):
        tok = self.next()
        if tok.kind != "id":
            raise SyntaxError(f"line {tok.line}: expected identifier, found {tok.text!r}")
        return tok.text

    # ---- module ----
    def module(self):
        self.expect("module")
        name = self.ident()
        ports = []
        self.expect("(")
        while not self.accept(")"):
            ports.append(self.ident())
            self.accept(",")
        self.expect(";")
        decls, blocks = [], []
        while not self.accept("endmodule"):
            tok = self.peek()
            if tok.text in ("input", "output", "reg", "integer"):
                decls += self.decl()
            elif tok.text == "parameter":
                self.next()
                n = self.ident()
                self.expect("=")
                decls.append(("param", n, self.expr(), tok.line))
                self.expect(";")
            elif tok.text == "always":
                blocks.append(self.always())
            else:
                raise SyntaxError(f"line {tok.line}: unsupported module item {tok.text!r}")
        return {"name": name, "ports": ports, "decls": decls, "blocks": blocks}

    def range_(self):
        if not self.accept("["):
            return None
        msb = self.expr()
        self.expect(":")
        lsb = self.expr()
        self.expect("]")
        return (msb, lsb)

    def decl(self):
        kind = self.next()
        rng = self.range_() if kind.text != "integer" else None
        out = []
        while True:
            n = self.ident()
            arr = self.range_()
            out.append((kind.text, n, rng, arr, kind.line))
            if self.accept(";"):
                break
            self.expect(",")
        return out

    def always(self):
        tok = self.expect("always")
        self.expect("@")
        self.expect("(")
        if self.accept("posedge"):
            clk = self.ident()
            self.expect(")")
            return ("seq", clk, self.stmt(), tok.line)
        self.expect("*")
        self.expect(")")
        return ("comb", None, self.stmt(), tok.line)

    # ---- statements ----
    def stmt(self):
        tok = self.peek()
        if self.accept("begin"):
            body = []
            while not self.accept("end"):
                body.append(self.stmt())
            return ("block", body)
        if self.accept("if"):
            self.expect("(")
            cond = self.expr()
            self.expect(")")
            then = self.stmt()
            other = self.stmt() if self.accept("else") else None
            return ("if", cond, then, other, tok.line)
        lv = self.lvalue()
        op = self.next()
        if op.text not in ("<=", "="):
            raise SyntaxError(f"line {op.line}: expected assignment, found {op.text!r}")
        rhs = self.expr()
        self.expect(";")
        return ("nba" if op.text == "<=" else "ba", lv, rhs, tok.line)

    def lvalue(self):
        n = self.ident()
        if self.accept("["):
            a = self.expr()
            if self.accept(":"):
                b = self.expr()
                self.expect("]")
                return ("part", n, a, b)
            self.expect("]")
            return ("index", n, a)
        return ("name", n)

    # ---- expressions (precedence climbing) ----
    LEVELS = [["||"], ["&&"], ["==", "!="], ["<", "<=", ">", ">="], ["<<", ">>"], ["+", "-"]]

    def expr(self, level=0):
        if level == len(self.LEVELS):
            return self.unary()
        lhs = self.expr(level + 1)
        while self.peek().kind == "op" and self.peek().text in self.LEVELS[level]:
            op = self.next().text
            rhs = self.expr(level + 1)
            lhs = ("bin", op, lhs, rhs)
        return lhs

    def unary(self):
        tok = self.peek()
        if tok.kind == "op" and tok.text in ("!", "~", "-"):
            self.next()
            return ("un", tok.text, self.unary())
        return self.primary()

    def primary(self):
        tok = self.next()
        if tok.kind == "num":
            return ("num", int(tok.text.replace("_", "")), None, 0)
        if tok.kind == "based":
            return self.based(tok)
        if tok.text == "(":
            e = self.expr()
            self.expect(")")
            return e
        if tok.kind == "id":
            node = ("name", tok.text)
            while self.accept("["):
                a = self.expr()
                if self.accept(":"):
                    b = self.expr()
                    self.expect("]")
                    node = ("partsel", node, a, b)
                else:
                    self.expect("]")
                    node = ("bitsel", node, a)
            return node
        raise SyntaxError(f"line {tok.line}: unexpected {tok.text!r} in expression")

    @staticmethod
    def based(tok):
        m = re.match(r"(?:(\d+)\s*)?'[sS]?([bBdDhHoO])\s*([0-9a-fA-FxXzZ_?]+)", tok.text)
        size = int(m.group(1)) if m.group(1) else 32
        base = {"b": 2, "d": 10, "h": 16, "o": 8}[m.group(2).lower()]
        digits = m.group(3).replace("_", "").lower()
        bits_per = {2: 1, 8: 3, 16: 4}.get(base)
        value, xmask = 0, 0
        if base == 10:
            if any(ch in "xz?" for ch in digits):
                xmask, value = (1 << size) - 1, 0
            else:
                value = int(digits)
        else:
            for ch in digits:
                value <<= bits_per
                xmask <<= bits_per
                if ch in "xz?":
                    xmask |= (1 << bits_per) - 1
                else:
                    value |= int(ch, base)
            # IEEE 1364 3.5.1: a literal whose leftmost digit is x/z is extended with x/z up to its size
            nbits = len(digits) * bits_per
            if digits[0] in "xz?" and size > nbits:
                xmask |= ((1 << size) - 1) & ~((1 << nbits) - 1)
        mask = (1 << size) - 1
        return ("num", value & mask, size, xmask & mask)


# ------------------------------------------------------------------------------------------------
# elaboration + C emission
# ------------------------------------------------------------------------------------------------
class Gen:
    def __init__(self, mod, runtime_params, src_name):
        self.m = mod
        self.rt = set(runtime_params)
        self.src = src_name
        self.params = {}          # folded constants
        self.sym = {}             # name -> dict(kind: scalar|wide|mem, width (int or C expr), depth, dir)
        self.order = []
        self.nba_written, self.ba_written = set(), set()
        self.elaborate()

    # ---- constant folding ----
    def const(self, e):
        """int value of a constant expression, or None if it depends on a runtime parameter / signal."""
        k = e[0]
        if k == "num":
            return None if e[3] else e[1]
        if k == "name":
            return self.params.get(e[1])
        if k == "un":
            v = self.const(e[2])
            if v is None:
                return None
            return {"-": -v, "!": int(not v), "~": ~v}[e[1]]
        if k == "bin":
            a, b = self.const(e[2]), self.const(e[3])
            if a is None or b is None:
                return None
            return {"+": a + b, "-": a - b, "<<": a << b, ">>": a >> b, "==": int(a == b), "!=": int(a != b),
                    "<": int(a < b), "<=": int(a <= b), ">": int(a > b), ">=": int(a >= b),
                    "&&": int(bool(a) and bool(b)), "||": int(bool(a) or bool(b))}[e[1]]
        return None

    def cexpr_const(self, e):
        """C expression for a width / bound that may depend on runtime parameters."""
        v = self.const(e)
        if v is not None:
            return str(v)
        k = e[0]
        if k == "name" and e[1] in self.rt:
            return f"((int64_t)m->P_{e[1]})"
        if k == "bin" and e[1] in ("+", "-"):
            return f"({self.cexpr_const(e[2])} {e[1]} {self.cexpr_const(e[3])})"
        raise SyntaxError(f"unsupported parameter expression {e!r}")

    def elaborate(self):
        for d in self.m["decls"]:
            if d[0] == "param":
                _, n, e, line = d
                if n in self.rt:
                    self.params.pop(n, None)
                    continue
                v = self.const(e)
                if v is None:
                    raise SyntaxError(f"line {line}: parameter {n} is not constant")
                self.params[n] = v
                continue
            kind, n, rng, arr, line = d
            if kind == "integer":
                ent = {"kind": "scalar", "width": 32, "dir": None}
            else:
                if rng is None:
                    w_const, w_c, lsb = 1, "1", 0
                else:
                    lsb = self.const(rng[1])
                    if lsb != 0:
                        raise SyntaxError(f"line {line}: {n}: only [msb:0] ranges are supported")
                    msb = self.const(rng[0])
                    w_const = None if msb is None else msb + 1
                    w_c = f"({self.cexpr_const(rng[0])} + 1)"
                if arr is not None:
                    hi, lo = self.const(arr[0]), self.const(arr[1])
                    if hi is None or lo != 0 or w_const is None or w_const > 64:
                        raise SyntaxError(f"line {line}: {n}: unsupported memory shape")
                    ent = {"kind": "mem", "width": w_const, "depth": hi + 1}
                elif w_const is not None and w_const <= 64:
                    ent = {"kind": "scalar", "width": w_const}
                else:
                    ent = {"kind": "wide", "width": w_c}
                ent["dir"] = kind if kind in ("input", "output") else None
            if n in self.sym:      # `output x; reg x;` pairs: the reg refines the port
                old = self.sym[n]
                ent["dir"] = old.get("dir") or ent.get("dir")
                if old["kind"] != ent["kind"] or str(old["width"]) != str(ent["width"]):
                    raise SyntaxError(f"line {line}: {n} redeclared with a different shape")
            else:
                self.order.append(n)
            self.sym[n] = ent
        for b in self.m["blocks"]:
            self.collect_writes(b[2], b[0])
        both = self.nba_written & self.ba_written
        if both:
            raise SyntaxError(f"assigned both blocking and non-blocking: {sorted(both)}")
        for n in self.nba_written | self.ba_written:
            if self.sym[n].get("dir") == "input":
                raise SyntaxError(f"input {n} is assigned")

    def collect_writes(self, s, blk):
        if s[0] == "block":
            for x in s[1]:
                self.collect_writes(x, blk)
        elif s[0] == "if":
            self.collect_writes(s[2], blk)
            if s[3]:
                self.collect_writes(s[3], blk)
        else:
            n = s[1][1]
            if n not in self.sym:
                raise SyntaxError(f"line {s[3]}: assignment to undeclared {n}")
            if (s[0] == "nba") != (blk == "seq"):
                raise SyntaxError(f"line {s[3]}: {'blocking' if s[0] == 'ba' else 'non-blocking'} assignment in a "
                                  f"{'clocked' if blk == 'seq' else 'combinational'} block is outside the subset")
            (self.nba_written if s[0] == "nba" else self.ba_written).add(n)

    # ---- expression emission: returns a C expression of type uint64_t ----
    def mask(self, w):
        return "0xFFFFFFFFFFFFFFFFull" if w >= 64 else hex((1 << w) - 1) + "ull"

    def rd(self, n):
        """C lvalue holding the CURRENT (pre-edge) value of scalar / mem n."""
        return f"m->cur.{n}" if n in self.nba_written else f"m->{n}"

    def ex(self, e):
        k = e[0]
        if k == "num":
            _, v, size, xmask = e
            if xmask:
                return f"({hex(v)}ull | (m->xfill ? {hex(xmask)}ull : 0ull))"
            return f"{v}ull" if v < (1 << 63) else f"{hex(v)}ull"
        if k == "name":
            n = e[1]
            if n in self.params:
                return f"{self.params[n]}ull"
            if n in self.rt:
                return f"m->P_{n}"
            ent = self.sym.get(n)
            if ent is None:
                raise SyntaxError(f"undeclared identifier {n}")
            if ent["kind"] != "scalar":
                raise SyntaxError(f"{n}: whole-vector / whole-memory reads only as the source of a whole-vector assignment")
            return self.rd(n)
        if k == "bitsel":
            base, idx = e[1], e[2]
            if base[0] == "name":
                n = base[1]
                ent = self.sym[n]
                if ent["kind"] == "mem":
                    return f"vs_mem_rd({self.rd(n)}, {ent['depth']}, {self.ex(idx)}, m->xfill, {self.mask(ent['width'])})"
                if ent["kind"] == "wide":
                    return f"vs_wide_bit(m->{n}, m->W_{n}, {self.ex(idx)})"
                return f"(({self.rd(n)} >> ({self.ex(idx)} & 63)) & 1ull)"   # scalar bit-select; idx < width in the subset's uses
            raise SyntaxError("bit-select of a select is outside the subset")
        if k == "partsel":
            base = e[1]
            msb, lsb = self.const(e[2]), self.const(e[3])
            if msb is None or lsb is None or msb < lsb or msb - lsb + 1 > 64:
                raise SyntaxError(f"part-select bounds must be constants spanning <= 64 bits: {e!r}")
            w = msb - lsb + 1
            if base[0] == "name":
                n = base[1]
                ent = self.sym[n]
                if ent["kind"] == "wide":
                    return f"vs_wide_get(m->{n}, m->W_{n}, {lsb}, {w})"
                if ent["kind"] == "scalar":
                    return f"(({self.rd(n)} >> {lsb}) & {self.mask(w)})"
                raise SyntaxError("part-select of a whole memory")
            if base[0] == "bitsel" and base[1][0] == "name" and self.sym[base[1][1]]["kind"] == "mem":
                return f"(({self.ex(base)} >> {lsb}) & {self.mask(w)})"
            raise SyntaxError(f"unsupported part-select base {base!r}")
        if k == "un":
            a = self.ex(e[2])
            if e[1] == "!":
                return f"((uint64_t)(({a}) == 0ull))"
            if e[1] == "~":
                return f"(~({a}))"
            return f"(0ull - ({a}))"
        if k == "bin":
            op, a, b = e[1], self.ex(e[2]), self.ex(e[3])
            if op in ("+", "-"):
                return f"(({a}) {op} ({b}))"
            if op in ("<<", ">>"):
                return f"vs_sh{'l' if op == '<<' else 'r'}({a}, {b})"
            if op in ("==", "!=", "<", "<=", ">", ">="):
                return f"((uint64_t)(({a}) {op} ({b})))"
            if op in ("&&", "||"):
                return f"((uint64_t)((({a}) != 0ull) {op} (({b}) != 0ull)))"
        raise SyntaxError(f"unsupported expression {e!r}")

    # ---- statements ----
    def st(self, s, ind, out):
        pad = "    " * ind
        if s[0] == "block":
            for x in s[1]:
                self.st(x, ind, out)
            return
        if s[0] == "if":
            out.append(f"{pad}if ({self.ex(s[1])}) {{   /* {self.src}:{s[4]} */")
            self.st(s[2], ind + 1, out)
            if s[3] is not None:
                out.append(f"{pad}}} else {{")
                self.st(s[3], ind + 1, out)
            out.append(f"{pad}}}")
            return
        kind, lv, rhs, line = s
        n = lv[1]
        ent = self.sym[n]
        nb = kind == "nba"
        tag = f"   /* {self.src}:{line} */"
        if ent["kind"] == "scalar":
            dst = f"m->nxt.{n}" if nb else f"m->{n}"
            if lv[0] == "name":
                out.append(f"{pad}{dst} = ({self.ex(rhs)}) & {self.mask(ent['width'])};{tag}")
            elif lv[0] == "index":
                out.append(f"{pad}{dst} = vs_setbit({dst}, {self.ex(lv[2])}, {ent['width']}, {self.ex(rhs)});{tag}")
            else:
                raise SyntaxError(f"line {line}: part-select assignment to scalar {n}")
        elif ent["kind"] == "mem":
            if lv[0] != "index":
                raise SyntaxError(f"line {line}: memory {n} must be assigned one word at a time")
            dst = f"m->nxt.{n}" if nb else f"m->{n}"
            out.append(f"{pad}vs_mem_wr({dst}, {ent['depth']}, {self.ex(lv[2])}, ({self.ex(rhs)}) & {self.mask(ent['width'])});{tag}")
        else:   # wide vector
            if not nb:
                raise SyntaxError(f"line {line}: blocking assignment to wide vector {n}")
            vid = self.wide_ids[n]
            if lv[0] == "index":
                out.append(f"{pad}vs_q_bit(m, {vid}, {self.ex(lv[2])}, {self.ex(rhs)});{tag}")
            elif lv[0] == "part":
                out.append(f"{pad}vs_q_part(m, {vid}, {self.cexpr_const(lv[2])}, {self.cexpr_const(lv[3])}, {self.ex(rhs)});{tag}")
            elif rhs[0] == "name" and rhs[1] in self.sym and self.sym[rhs[1]]["kind"] == "wide":
                out.append(f"{pad}vs_q_copy(m, {vid}, {self.wide_ids[rhs[1]]});{tag}")
            else:
                out.append(f"{pad}vs_q_part(m, {vid}, (int64_t)m->W_{n} - 1, 0, {self.ex(rhs)});{tag}")

    # ---- whole file ----
    def emit(self):
        name = self.m["name"]
        wides = [n for n in self.order if self.sym[n]["kind"] == "wide"]
        self.wide_ids = {n: i for i, n in enumerate(wides)}
        seq_regs = [n for n in self.order if n in self.nba_written and self.sym[n]["kind"] != "wide"]
        other = [n for n in self.order if n not in self.nba_written and self.sym[n]["kind"] != "wide"]

        def field(n):
            ent = self.sym[n]
            return f"uint64_t {n}[{ent['depth']}];" if ent["kind"] == "mem" else f"uint64_t {n};"

        o = []
        o.append(f"/* GENERATED by oracle/vsim/v2c.py from {self.src} -- do not edit, do not commit.")
        o.append(" * TEST INFRASTRUCTURE ONLY: C model of the reference's own Verilog (see oracle/vsim/README.md). */")
        o.append('#include "vsim_rt.h"')
        o.append("typedef struct {")
        for n in seq_regs:
            o.append(f"    {field(n)}")
        if not seq_regs:
            o.append("    uint64_t unused_;")
        o.append("} vs_regs;")
        o.append("struct vs_model {")
        o.append("    VS_MODEL_HEADER")
        for n in sorted(self.rt):
            o.append(f"    uint64_t P_{n};")
        o.append("    vs_regs cur, nxt;      /* registers assigned by non-blocking assignments: pre-edge / post-edge */")
        for n in other:
            o.append(f"    {field(n)}   /* {'input' if self.sym[n].get('dir') == 'input' else 'blocking-assigned / unassigned'} */")
        for n in wides:
            o.append(f"    uint64_t *{n}; uint64_t W_{n};")
        o.append("};")
        o.append("")
        o.append(f"const char *vs_module_name(void) {{ return \"{name}\"; }}")
        o.append(f"int vs_n_wide(void) {{ return {len(wides)}; }}")
        o.append("")
        # constructor
        o.append("vs_model *vs_new(const char *const *pnames, const uint64_t *pvalues, int np, int xfill) {")
        o.append("    vs_model *m = (vs_model *)calloc(1, sizeof(vs_model));")
        o.append("    if (!m) return NULL;")
        o.append("    m->xfill = xfill ? 1 : 0;")
        for n in sorted(self.rt):
            o.append(f"    {{ int found = 0; for (int k = 0; k < np; k++) if (!strcmp(pnames[k], \"{n}\")) {{ m->P_{n} = pvalues[k]; found = 1; }}")
            o.append("      if (!found) { free(m); return NULL; } }")
        o.append("    const uint64_t fill = xfill ? ~0ull : 0ull;   /* power-up value of every reg: x */")
        for n in seq_regs + other:
            ent = self.sym[n]
            if ent["kind"] == "mem":
                o.append(f"    for (int k = 0; k < {ent['depth']}; k++) m->{'cur.' if n in self.nba_written else ''}{n}[k] = fill & {self.mask(ent['width'])};")
            else:
                o.append(f"    m->{'cur.' if n in self.nba_written else ''}{n} = fill & {self.mask(ent['width'])};")
        o.append(f"    m->n_wide = {len(wides)};")
        for n in wides:
            o.append(f"    m->W_{n} = (uint64_t)({self.sym[n]['width']});")
            o.append(f"    m->{n} = vs_wide_new(m->W_{n}, xfill);")
            o.append(f"    m->wide[{self.wide_ids[n]}] = m->{n}; m->wide_w[{self.wide_ids[n]}] = m->W_{n};")
        o.append("    return m;")
        o.append("}")
        o.append("void vs_delete(vs_model *m) { if (!m) return; for (int k = 0; k < m->n_wide; k++) free(m->wide[k]); free(m->q); free(m->snap); free(m); }")
        o.append("")
        # port access
        o.append("int vs_set(vs_model *m, const char *port, uint64_t v) {")
        for n in self.order:
            ent = self.sym[n]
            if ent.get("dir") == "input" and ent["kind"] == "scalar":
                o.append(f"    if (!strcmp(port, \"{n}\")) {{ m->{n} = v & {self.mask(ent['width'])}; return 0; }}")
        o.append("    return -1;")
        o.append("}")
        o.append("int vs_set_wide(vs_model *m, const char *port, const uint64_t *words, uint64_t nbits) {")
        for n in wides:
            if self.sym[n].get("dir") == "input":
                o.append(f"    if (!strcmp(port, \"{n}\")) {{ vs_wide_load(m->{n}, m->W_{n}, words, nbits, m->xfill); return 0; }}")
        o.append("    return -1;")
        o.append("}")
        o.append("int vs_get(const vs_model *m, const char *name, uint64_t *v) {")
        for n in self.order:
            if self.sym[n]["kind"] == "scalar":
                o.append(f"    if (!strcmp(name, \"{n}\")) {{ *v = {self.rd(n)}; return 0; }}")
        o.append("    return -1;")
        o.append("}")
        o.append("uint64_t *vs_ptr(vs_model *m, const char *name) {")
        for n in self.order:
            if self.sym[n]["kind"] == "scalar":
                o.append(f"    if (!strcmp(name, \"{n}\")) return &{self.rd(n)};")
        o.append("    return NULL;")
        o.append("}")
        o.append("uint64_t *vs_get_wide(vs_model *m, const char *name, uint64_t *nbits) {")
        for n in wides:
            o.append(f"    if (!strcmp(name, \"{n}\")) {{ *nbits = m->W_{n}; return m->{n}; }}")
        o.append("    return NULL;")
        o.append("}")
        o.append("")
        # combinational blocks
        comb = [b for b in self.m["blocks"] if b[0] == "comb"]
        seq = [b for b in self.m["blocks"] if b[0] == "seq"]
        o.append("/* always @(*) blocks, in source order */")
        o.append("void vs_eval_comb(vs_model *m) {")
        for b in comb:
            o.append(f"    /* {self.src}:{b[3]} */")
            self.st(b[2], 1, o)
        o.append("}")
        o.append("")
        clocks = sorted({b[1] for b in seq})
        if len(clocks) != 1:
            raise SyntaxError(f"exactly one clock expected, found {clocks}")
        o.append(f"/* always @(posedge {clocks[0]}): evaluate (reads pre-edge values), then commit */")
        o.append("void vs_posedge(vs_model *m) {")
        o.append("    vs_eval_comb(m);")
        o.append("    m->nxt = m->cur;")
        o.append("    m->qn = 0;")
        for b in seq:
            o.append(f"    /* {self.src}:{b[3]} */")
            self.st(b[2], 1, o)
        o.append("    m->cur = m->nxt;")
        o.append("    vs_q_commit(m);")
        o.append("}")
        return "\n".join(o) + "\n"


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("source")
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--runtime-param", action="append", default=[])
    args = ap.parse_args()
    src = open(args.source).read()
    mod = Parser(lex(src)).module()
    gen = Gen(mod, args.runtime_param, args.source)
    text = gen.emit()
    with open(args.output, "w") as f:
        f.write(text)
    print(f"v2c: {args.source}: module {mod['name']}, {len(mod['decls'])} declarations, {len(mod['blocks'])} always blocks "
          f"-> {args.output} ({len(text.splitlines())} lines)", file=sys.stderr)


if __name__ == "__main__":
    main()
