"""glsl_simt.py -- a GLSL-subset interpreter that EXECUTES the reference's compute shaders.

Test infrastructure.  It reads the reference's own, unmodified `*_comp.glsl` text (from
/root/reference at fixture-generation time -- nothing is copied into this repo), parses the
subset of GLSL 4.40 those three files use, and runs `main()` for all invocations of a dispatch
in lock step (SIMT: one numpy lane per invocation, an execution mask for divergent `if`,
`continue` and `return`).  Arithmetic is IEEE fp32, one rounding per GLSL operation in the
order the GLSL grammar gives (left-associative, no contraction).

Built-ins whose precision GLSL leaves to the driver are pinned like this:
    pow(x, y)     correctly rounded: evaluated in float64, rounded once to fp32
    length(v)     sqrt(x*x + y*y + z*z), fp32, correctly rounded sqrt
    normalize(v)  v / length(v)  (IEEE divide; NaN for the zero vector)
    max(a, b)     (a < b) ? b : a   (the GLSL specification's definition)

Supported: #define (object-like), struct, std430 buffer blocks with an unsized array member,
std140 uniform blocks, `const` globals, uniform declarations, `void main()`, declarations,
assignments (= += -= *= /=) to variables / swizzles / buffer members, if / else if / else,
for loops with a uniform trip count, `continue`, `return`, ++/--, arithmetic / comparison /
logical operators, constructors vec3/vec4/float/uint/int, swizzles and array indexing.
"""
from __future__ import annotations

import re

import numpy as np

F32 = np.float32
SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}
VEC_SIZES = {"vec2": 2, "vec3": 3, "vec4": 4}
TYPES = {"float", "int", "uint", "bool", "vec2", "vec3", "vec4", "mat4", "void"}


# ---------------------------------------------------------------------------------------------
# lexer / preprocessor
# ---------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fFuU]?)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>\+\+|--|\+=|-=|\*=|/=|<=|>=|==|!=|&&|\|\||[-+*/<>=!(){}\[\];,.?:])
  | (?P<ws>\s+)
""", re.X)


def preprocess(src: str, defines: dict | None = None):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    macros = {}
    lines = []
    for line in src.split("\n"):
        s = line.strip()
        if s.startswith("#define"):
            parts = s.split(None, 2)
            macros[parts[1]] = parts[2] if len(parts) > 2 else ""
        elif s.startswith("#"):
            continue
        else:
            lines.append(line)
    if defines:
        for k, v in defines.items():
            if k not in macros:
                raise KeyError(f"shader has no #define {k}")
            macros[k] = str(v)
    return "\n".join(lines), macros


def tokenize(text: str, macros: dict):
    out = []
    pos = 0
    while pos < len(text):
        m = TOKEN_RE.match(text, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {text[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        tok = m.group()
        if kind == "id" and tok in macros:
            out.extend(tokenize(macros[tok], macros))
        else:
            out.append((kind, tok))
    return out


# ---------------------------------------------------------------------------------------------
# parser -> tuples
# ---------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0
        self.structs, self.buffers, self.uniform_members = {}, {}, {}
        self.globals, self.functions = [], {}

    def peek(self, k=0):
        return self.t[self.i + k][1] if self.i + k < len(self.t) else None

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expect(self, s):
        kind, tok = self.next()
        if tok != s:
            raise SyntaxError(f"expected {s!r}, got {tok!r}")

    def accept(self, s):
        if self.peek() == s:
            self.i += 1
            return True
        return False

    # -- top level ---------------------------------------------------------------------------
    def parse_unit(self):
        while self.i < len(self.t):
            self.parse_external()
        return self

    def skip_layout(self):
        quals = []
        if self.accept("layout"):
            self.expect("(")
            depth = 1
            while depth:
                tok = self.next()[1]
                depth += tok == "("
                depth -= tok == ")"
                quals.append(tok)
        return quals

    def parse_member_list(self):
        members = []
        self.expect("{")
        while not self.accept("}"):
            ty = self.next()[1]
            name = self.next()[1]
            arr = None
            if self.accept("["):
                arr = -1
                if self.peek() != "]":
                    arr = int(self.next()[1])
                self.expect("]")
            self.expect(";")
            members.append((ty, name, arr))
        return members

    def parse_external(self):
        self.skip_layout()
        tok = self.peek()
        if tok == "in":
            self.next(); self.expect(";")
        elif tok == "struct":
            self.next()
            name = self.next()[1]
            self.structs[name] = self.parse_member_list()
            self.expect(";")
        elif tok == "buffer":
            self.next()
            self.next()                                   # block name
            for ty, name, arr in self.parse_member_list():
                self.buffers[name] = ty
            if self.peek() != ";":
                self.next()
            self.expect(";")
        elif tok == "uniform":
            self.next()
            if self.peek(1) == "{":
                self.next()
                for ty, name, arr in self.parse_member_list():
                    self.uniform_members[name] = ty
                self.expect(";")
            else:
                ty = self.next()[1]
                name = self.next()[1]
                self.expect(";")
                self.uniform_members[name] = ty
        elif tok == "const":
            self.globals.append(self.parse_statement())
        elif tok in TYPES and self.peek(2) == "(":
            self.next()
            name = self.next()[1]
            self.expect("("); self.expect(")")
            self.functions[name] = self.parse_block()
        else:
            raise SyntaxError(f"unsupported top-level construct at {tok!r}")

    # -- statements ---------------------------------------------------------------------------
    def parse_block(self):
        self.expect("{")
        body = []
        while not self.accept("}"):
            body.append(self.parse_statement())
        return ("block", body)

    def parse_statement(self):
        tok = self.peek()
        if tok == "{":
            return self.parse_block()
        if tok == "if":
            self.next(); self.expect("(")
            cond = self.parse_expr(); self.expect(")")
            then = self.parse_statement()
            other = None
            if self.accept("else"):
                other = self.parse_statement()
            return ("if", cond, then, other)
        if tok == "for":
            self.next(); self.expect("(")
            init = self.parse_statement()
            cond = self.parse_expr(); self.expect(";")
            incr = self.parse_expr(); self.expect(")")
            return ("for", init, cond, incr, self.parse_statement())
        if tok == "return":
            self.next(); self.expect(";")
            return ("return",)
        if tok == "continue":
            self.next(); self.expect(";")
            return ("continue",)
        if tok == "const" or tok in TYPES:
            self.accept("const")
            ty = self.next()[1]
            name = self.next()[1]
            init = None
            if self.accept("="):
                init = self.parse_expr()
            self.expect(";")
            return ("decl", ty, name, init)
        e = self.parse_expr()
        self.expect(";")
        return ("expr", e)

    # -- expressions (precedence climbing) --------------------------------------------------------
    BIN = [("||",), ("&&",), ("==", "!="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/")]

    def parse_expr(self):
        lhs = self.parse_binary(0)
        if self.peek() in ("=", "+=", "-=", "*=", "/="):
            op = self.next()[1]
            return ("assign", op, lhs, self.parse_expr())
        return lhs

    def parse_binary(self, level):
        if level == len(self.BIN):
            return self.parse_unary()
        lhs = self.parse_binary(level + 1)
        while self.peek() in self.BIN[level]:
            op = self.next()[1]
            lhs = ("bin", op, lhs, self.parse_binary(level + 1))
        return lhs

    def parse_unary(self):
        if self.peek() in ("-", "!", "+"):
            op = self.next()[1]
            return ("un", op, self.parse_unary())
        return self.parse_postfix()

    def parse_postfix(self):
        kind, tok = self.next()
        if kind == "num":
            e = ("num", tok)
        elif tok == "(":
            e = self.parse_expr(); self.expect(")")
        elif kind == "id":
            if self.peek() == "(":
                self.next()
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.parse_expr())
                        if self.accept(")"):
                            break
                        self.expect(",")
                e = ("call", tok, args)
            else:
                e = ("var", tok)
        else:
            raise SyntaxError(f"unexpected token {tok!r}")
        while True:
            if self.accept("."):
                e = ("member", e, self.next()[1])
            elif self.accept("["):
                idx = self.parse_expr(); self.expect("]")
                e = ("index", e, idx)
            elif self.peek() in ("++", "--"):
                e = ("postinc", self.next()[1], e)
            else:
                return e


# ---------------------------------------------------------------------------------------------
# SIMT evaluator
# ---------------------------------------------------------------------------------------------
class Vec:
    """GLSL vecN: a list of components, each a scalar or a per-lane array."""
    __slots__ = ("c",)

    def __init__(self, comps):
        self.c = list(comps)


class BufRef:
    """particles[idx] / particles[idx].field -- resolved lazily so it can be loaded or stored."""
    __slots__ = ("buf", "idx", "field")

    def __init__(self, buf, idx, field=None):
        self.buf, self.idx, self.field = buf, idx, field


def _f32(x):
    if isinstance(x, np.ndarray):
        return x.astype(F32, copy=False) if x.dtype != F32 else x
    return F32(x)


def _is_int(x):
    if isinstance(x, (bool, np.bool_)):
        return False
    if isinstance(x, (int, np.integer)):
        return True
    return isinstance(x, np.ndarray) and x.dtype.kind in "iu"


def _arith(op, a, b):
    if isinstance(a, Vec) or isinstance(b, Vec):
        n = len(a.c) if isinstance(a, Vec) else len(b.c)
        ac = a.c if isinstance(a, Vec) else [a] * n
        bc = b.c if isinstance(b, Vec) else [b] * n
        return Vec([_arith(op, x, y) for x, y in zip(ac, bc)])
    if not (_is_int(a) and _is_int(b)):          # implicit int -> float conversion
        a, b = _f32(a), _f32(b)
    with np.errstate(all="ignore"):
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            return a // b if _is_int(a) else a / b
    raise ValueError(op)


def _compare(op, a, b):
    if not (_is_int(a) and _is_int(b)):
        a, b = _f32(a), _f32(b)
    with np.errstate(invalid="ignore"):
        return {"<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal,
                "==": np.equal, "!=": np.not_equal}[op](a, b)


def _pow(x, y):
    with np.errstate(all="ignore"):
        return np.power(np.asarray(x, np.float64), np.float64(y)).astype(F32)


def _length(v):
    acc = None
    for comp in v.c:
        sq = _arith("*", comp, comp)
        acc = sq if acc is None else _arith("+", acc, sq)
    with np.errstate(invalid="ignore"):
        return np.sqrt(_f32(acc))


class Shader:
    def __init__(self, source: str, defines: dict | None = None):
        text, self.macros = preprocess(source, defines)
        self.ast = Parser(tokenize(text, self.macros)).parse_unit()
        self.local_size = None
        m = re.search(r"local_size_x\s*=\s*(\w+)", text)
        if m:
            v = m.group(1)
            self.local_size = int(self.macros.get(v, v))

    # -- running a dispatch --------------------------------------------------------------------
    def dispatch(self, num_groups: int, buffers: dict, uniforms: dict):
        """buffers: name -> {field: float32 array (n, 4)}; uniforms: name -> float / sequence."""
        lanes = num_groups * self.local_size
        self.n = lanes
        self.buffers = buffers
        self.scopes = [{}]
        g = self.scopes[0]
        for name, ty in self.ast.uniform_members.items():
            if name in uniforms:
                val = uniforms[name]
                g[name] = Vec([F32(c) for c in val]) if ty in VEC_SIZES else F32(val)
        for name in self.ast.buffers:
            g[name] = ("buffer", name)
        g["gl_GlobalInvocationID"] = Vec([np.arange(lanes, dtype=np.uint32), np.uint32(0), np.uint32(0)])
        self.mask = np.ones(lanes, bool)
        self.returned = np.zeros(lanes, bool)
        self.cont = None
        for st in self.ast.globals:
            self.exec(st)
        self.scopes.append({})
        self.exec(self.ast.functions["main"])

    # -- helpers -------------------------------------------------------------------------------
    def lookup(self, name):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc, sc[name]
        raise NameError(f"undeclared identifier {name!r} (uniform not supplied?)")

    def active(self):
        m = self.mask & ~self.returned
        if self.cont is not None:
            m = m & ~self.cont
        return m

    def blend(self, old, new, mask):
        """Masked write of one scalar slot.  A uniform value written outside any divergent
        region (every live lane active) stays uniform -- e.g. the loop counter `j++`."""
        if not isinstance(old, np.ndarray) and not isinstance(new, np.ndarray):
            live = ~self.returned if self.cont is None else ~self.returned & ~self.cont
            if (mask == live).all():
                return new
        if mask.all():
            return new
        if not mask.any():
            return old
        return np.where(mask, new, old)

    def convert(self, ty, v):
        if ty == "float":
            return _f32(v)
        if ty in ("uint", "int"):
            dt = np.uint32 if ty == "uint" else np.int32
            return v.astype(dt) if isinstance(v, np.ndarray) else (int(v) if _is_int(v) else int(v))
        if ty in VEC_SIZES:
            assert isinstance(v, Vec) and len(v.c) == VEC_SIZES[ty], f"cannot convert to {ty}"
            return Vec([_f32(c) for c in v.c])
        return v

    # -- buffer access -----------------------------------------------------------------------------
    def buf_load(self, ref: BufRef):
        arr = self.buffers[ref.buf][ref.field]
        idx = ref.idx
        if isinstance(idx, np.ndarray):
            rows = arr[np.clip(idx.astype(np.int64), 0, len(arr) - 1)]
            return Vec([rows[:, k] for k in range(4)])
        row = arr[int(idx)]
        return Vec([row[k] for k in range(4)])

    def buf_store(self, ref: BufRef, comps, values):
        arr = self.buffers[ref.buf][ref.field]
        m = self.active()
        idx = ref.idx
        if not isinstance(idx, np.ndarray):
            raise NotImplementedError("stores through a uniform index")
        rows = idx[m].astype(np.int64)
        for k, val in zip(comps, values):
            val = _f32(val)
            arr[rows, k] = val[m] if isinstance(val, np.ndarray) else val

    # -- expression evaluation ------------------------------------------------------------------
    def ev(self, e):
        kind = e[0]
        if kind == "num":
            tok = e[1]
            if tok[-1] in "uU":
                return int(tok[:-1])
            if tok[-1] in "fF" or any(ch in tok for ch in ".eE"):
                return F32(tok.rstrip("fF"))
            return int(tok)
        if kind == "var":
            return self.lookup(e[1])[1]
        if kind == "un":
            v = self.ev(e[2])
            if e[1] == "-":
                return Vec([-c for c in v.c]) if isinstance(v, Vec) else -v
            if e[1] == "!":
                return np.logical_not(v)
            return v
        if kind == "bin":
            op = e[1]
            a, b = self.ev(e[2]), self.ev(e[3])
            if op in "+-*/":
                return _arith(op, a, b)
            if op in ("&&", "||"):
                return np.logical_and(a, b) if op == "&&" else np.logical_or(a, b)
            return _compare(op, a, b)
        if kind == "call":
            return self.call(e[1], [self.ev(a) for a in e[2]])
        if kind == "member":
            base = self.ev(e[1])
            if isinstance(base, BufRef):
                if base.field is None:
                    return BufRef(base.buf, base.idx, e[2])
                base = self.buf_load(base)
            comps = [base.c[SWZ[ch]] for ch in e[2]]
            return comps[0] if len(comps) == 1 else Vec(comps)
        if kind == "index":
            base = self.ev(e[1])
            idx = self.ev(e[2])
            if isinstance(base, tuple) and base[0] == "buffer":
                return BufRef(base[1], idx)
            if isinstance(base, BufRef):
                base = self.buf_load(base)
            return base.c[int(idx)]
        if kind == "assign":
            return self.assign(e[1], e[2], self.ev(e[3]))
        if kind == "postinc":
            old = self.ev(e[2])
            self.assign("=", e[2], old + 1 if e[1] == "++" else old - 1)
            return old
        raise NotImplementedError(kind)

    def call(self, name, args):
        if name in VEC_SIZES:
            comps = []
            for a in args:
                comps.extend(a.c if isinstance(a, Vec) else [a])
            n = VEC_SIZES[name]
            if len(comps) == 1:
                comps = comps * n
            assert len(comps) == n, f"{name}() with {len(comps)} components"
            return Vec([_f32(c) for c in comps])
        if name == "float":
            return _f32(args[0])
        if name in ("uint", "int"):
            return self.convert(name, args[0])
        if name == "length":
            return _length(args[0])
        if name == "normalize":
            ln = _length(args[0])
            return Vec([_arith("/", c, ln) for c in args[0].c])
        if name == "pow":
            x, y = args
            if isinstance(x, Vec):
                return Vec([_pow(c, y) for c in x.c])
            return _pow(x, y)
        if name == "max":
            a, b = _f32(args[0]), _f32(args[1])
            with np.errstate(invalid="ignore"):
                return np.where(a < b, b, a) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) \
                    else (b if a < b else a)
        if name == "dot":
            acc = None
            for x, y in zip(args[0].c, args[1].c):
                t = _arith("*", x, y)
                acc = t if acc is None else _arith("+", acc, t)
            return acc
        if name == "sqrt":
            return np.sqrt(_f32(args[0]))
        raise NotImplementedError(f"built-in {name}()")

    # -- assignment ------------------------------------------------------------------------------
    def assign(self, op, target, value):
        if op != "=":
            value = _arith(op[0], self.ev(target), value)
        m = self.active()
        kind = target[0]
        if kind == "var":
            scope, old = self.lookup(target[1])
            if isinstance(old, Vec):
                vc = value.c if isinstance(value, Vec) else [value] * len(old.c)
                scope[target[1]] = Vec([self.blend(o, _f32(v), m) for o, v in zip(old.c, vc)])
            else:
                if not _is_int(old):
                    value = _f32(value)
                scope[target[1]] = self.blend(old, value, m)
            return value
        if kind == "member":
            base = self.ev(target[1])
            sw = [SWZ[ch] for ch in target[2]]
            vals = value.c if isinstance(value, Vec) else [value] * len(sw)
            if isinstance(base, BufRef) and base.field is not None:
                self.buf_store(base, sw, vals)
                return value
            if target[1][0] == "var":                  # local_vec.x = ... / local_vec.xyz = ...
                scope, old = self.lookup(target[1][1])
                comps = list(old.c)
                for k, v in zip(sw, vals):
                    comps[k] = self.blend(comps[k], _f32(v), m)
                scope[target[1][1]] = Vec(comps)
                return value
        if kind == "index":
            base = self.ev(target[1])
            if isinstance(base, BufRef) and base.field is not None:
                self.buf_store(base, [int(self.ev(target[2]))], [value])
                return value
        raise NotImplementedError(f"assignment target {target}")

    # -- statements ---------------------------------------------------------------------------------
    def exec(self, st):
        kind = st[0]
        if kind == "block":
            self.scopes.append({})
            for s in st[1]:
                if not self.active().any():
                    break
                self.exec(s)
            self.scopes.pop()
        elif kind == "decl":
            _, ty, name, init = st
            if init is None:
                val = Vec([F32(0)] * VEC_SIZES[ty]) if ty in VEC_SIZES else (F32(0) if ty == "float" else 0)
            else:
                val = self.convert(ty, self.ev(init))
            self.scopes[-1][name] = val
        elif kind == "expr":
            self.ev(st[1])
        elif kind == "if":
            cond = self.ev(st[1])
            if isinstance(cond, np.ndarray) and cond.ndim:
                saved = self.mask
                self.mask = saved & cond
                if self.active().any():
                    self.exec(st[2])
                if st[3] is not None:
                    self.mask = saved & ~cond
                    if self.active().any():
                        self.exec(st[3])
                self.mask = saved
            elif cond:
                self.exec(st[2])
            elif st[3] is not None:
                self.exec(st[3])
        elif kind == "for":
            _, init, cond, incr, body = st
            self.scopes.append({})
            self.exec(init)
            outer_cont = self.cont
            while True:
                c = self.ev(cond)
                if isinstance(c, np.ndarray) and c.ndim:
                    raise NotImplementedError("for loop with a per-lane trip count")
                if not c or not (self.mask & ~self.returned).any():
                    break
                self.cont = np.zeros(self.n, bool)
                self.exec(body)
                self.cont = None
                self.ev(incr)
            self.cont = outer_cont
            self.scopes.pop()
        elif kind == "return":
            self.returned = self.returned | self.active()
        elif kind == "continue":
            if self.cont is None:
                raise SyntaxError("continue outside a loop")
            self.cont = self.cont | self.active()
        else:
            raise NotImplementedError(kind)
