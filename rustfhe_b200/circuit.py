"""Circuit front-end over the batched gate engine (BASELINE config 4).

The reference evaluates logic expressions depth-first, ONE bootstrapped gate at a time (`eval_logic_expr`,
nander/src/lib.rs:72-89).  Independent gates are where the GPU's throughput comes from, so this module levelises a gate
netlist and submits every level as batched calls (one `tfhe_b200_gate_batch*` per opcode per level).

Mirrored reference interface (nander/src/lib.rs):
  parse_logic_expr(str) -> LogicExpr        grammar `0 1 & | ^ ! $ ( )`, left-assoc, no precedence   lib.rs:90-172
  eval_logic_expr(pros, expr)               pros = TFHE (native gates, lib.rs:40-62)                    lib.rs:72-89
  Logip mapping for TFHE: nand/not/and/or/xor -> hom_nand/hom_not/hom_and/hom_or/hom_xor               lib.rs:40-62
Leaves are TRIVIAL ciphertexts TLWERep::logic_true/false (tlwe.rs:80-87), as in the reference.
"""
from dataclasses import dataclass, field

import numpy as np

from . import _capi as K

NAND, AND, OR, XOR, NOT = K.NAND, K.AND, K.OR, K.XOR, K.NOT
_SYM = {"&": AND, "|": OR, "^": XOR, "$": NAND}


# ------------------------------------------------------------------------------------------------------------
# nander grammar
# ------------------------------------------------------------------------------------------------------------
@dataclass
class LogicExpr:
    """LogicExpr<R> (nander/src/lib.rs:64-71): kind in {'nand','not','and','or','xor','leaf'}."""
    kind: str
    lhs: "LogicExpr" = None
    rhs: "LogicExpr" = None
    value: int = 0


def parse_logic_expr(text):
    """nander/src/lib.rs:90-172.  Whitespace is dropped; binary operators are left-associative without precedence; `!`
    binds to the following element; input after a complete expression is ignored, like the reference.
    Raises ValueError with the reference's messages."""
    s = [c for c in text.strip() if not c.isspace()]
    pos = 0

    def peek():
        return s[pos] if pos < len(s) else None

    def parse_elem():
        nonlocal pos
        c = peek()
        if c is None:
            raise ValueError("invalid element. this is none")
        pos += 1
        if c == "0":
            return LogicExpr("leaf", value=0)
        if c == "1":
            return LogicExpr("leaf", value=1)
        if c == "(":
            e = parse_binary()
            if peek() != ")":
                raise ValueError("braket is not closed")
            pos += 1
            return e
        raise ValueError("invalid element")

    def parse_mono():
        nonlocal pos
        if peek() == "!":
            pos += 1
            return LogicExpr("not", lhs=parse_mono())
        return parse_elem()

    def parse_binary():
        nonlocal pos
        lhs = parse_mono()
        while peek() in _SYM:
            op = {"&": "and", "|": "or", "^": "xor", "$": "nand"}[peek()]
            pos += 1
            lhs = LogicExpr(op, lhs=lhs, rhs=parse_mono())
        return lhs

    return parse_binary()


def eval_logic_expr_plain(expr):
    """Cleartext semantics of a LogicExpr (what the decrypted result must equal)."""
    k = expr.kind
    if k == "leaf":
        return expr.value
    a = eval_logic_expr_plain(expr.lhs)
    if k == "not":
        return 1 - a
    b = eval_logic_expr_plain(expr.rhs)
    return {"nand": 1 - (a & b), "and": a & b, "or": a | b, "xor": a ^ b}[k]


# ------------------------------------------------------------------------------------------------------------
# netlists
# ------------------------------------------------------------------------------------------------------------
@dataclass
class Netlist:
    """Gate netlist over wires 0..n_wires-1. Wires [0, n_inputs) are primary inputs; `consts` maps wire -> 0/1 (trivial
    ciphertexts).  gates: (op, in0, in1 or -1, out)."""
    n_inputs: int = 0
    n_wires: int = 0
    gates: list = field(default_factory=list)
    consts: dict = field(default_factory=dict)
    outputs: list = field(default_factory=list)

    def new_wire(self):
        self.n_wires += 1
        return self.n_wires - 1

    def add_inputs(self, k):
        assert self.n_wires == self.n_inputs, "declare inputs first"
        first = self.n_wires
        self.n_inputs += k
        self.n_wires += k
        return list(range(first, first + k))

    def const(self, bit):
        w = self.new_wire()
        self.consts[w] = int(bit)
        return w

    def gate(self, op, a, b=-1):
        out = self.new_wire()
        self.gates.append((op, a, b, out))
        return out

    def nand(self, a, b):
        return self.gate(NAND, a, b)

    def levels(self):
        """ASAP levelisation: level(g) = 1 + max(level of its inputs); inputs and constants are level 0.
        Returns a list of levels, each a dict op -> (in0 idx array, in1 idx array, out idx array)."""
        lvl = np.zeros(self.n_wires, np.int64)
        per = {}
        for op, a, b, out in self.gates:
            l = 1 + max(lvl[a], lvl[b] if b >= 0 else 0)
            lvl[out] = l
            per.setdefault(int(l), {}).setdefault(op, []).append((a, b if b >= 0 else a, out))
        res = []
        for l in sorted(per):
            res.append({op: tuple(np.array(col, np.int64) for col in zip(*g)) for op, g in per[l].items()})
        return res

    def simulate(self, input_bits):
        """Cleartext evaluation (oracle for the tests)."""
        v = np.zeros(self.n_wires, np.uint8)
        v[:self.n_inputs] = np.asarray(input_bits, np.uint8)
        for w, bit in self.consts.items():
            v[w] = bit
        for op, a, b, out in self.gates:
            x, y = int(v[a]), int(v[b]) if b >= 0 else 0
            v[out] = {NAND: 1 - (x & y), AND: x & y, OR: x | y, XOR: x ^ y, NOT: 1 - x}[op]
        return v[self.outputs] if self.outputs else v


def ripple_carry_adder(nbits=32):
    """NAND-only ripple-carry adder (BASELINE config 4): bit 0 = 5-NAND half adder, bits 1.. = 9-NAND full adders.
    Inputs: x[0..nbits), y[0..nbits) little endian; outputs: nbits sum bits + carry out."""
    nl = Netlist()
    x = nl.add_inputs(nbits)
    y = nl.add_inputs(nbits)
    t = nl.nand(x[0], y[0])
    s = nl.nand(nl.nand(x[0], t), nl.nand(t, y[0]))
    c = nl.nand(t, t)
    outs = [s]
    for i in range(1, nbits):
        x1 = nl.nand(x[i], y[i])
        s1 = nl.nand(nl.nand(x[i], x1), nl.nand(y[i], x1))   # x ^ y
        x4 = nl.nand(s1, c)
        outs.append(nl.nand(nl.nand(s1, x4), nl.nand(c, x4)))  # x ^ y ^ c
        c = nl.nand(x1, x4)                                    # majority
    outs.append(c)
    nl.outputs = outs
    return nl


def prefix_adder(nbits=32):
    """Kogge-Stone parallel-prefix adder over the TFHE engine's native gates (hom_and / hom_or / hom_xor, the Logip mapping of
    nander/src/lib.rs:40-62): 2 + 2 log2(nbits) levels of width ~nbits..3 nbits instead of the ripple-carry chain's 2 nbits
    levels of width <= 3 -- the shape on which level-synchronous batching turns gate THROUGHPUT into circuit latency.
    Same interface as ripple_carry_adder: inputs x[0..nbits), y[0..nbits) little endian; outputs nbits sum bits + carry out."""
    nl = Netlist()
    x = nl.add_inputs(nbits)
    y = nl.add_inputs(nbits)
    g = [nl.gate(AND, x[i], y[i]) for i in range(nbits)]      # generate
    p = [nl.gate(XOR, x[i], y[i]) for i in range(nbits)]      # propagate (also the half sum)
    G, P = list(g), list(p)
    d = 1
    while d < nbits:
        G2, P2 = list(G), list(P)
        for i in range(d, nbits):
            t = nl.gate(AND, P[i], G[i - d])
            G2[i] = nl.gate(OR, G[i], t)                      # G[i] | (P[i] & G[i-d])
            if i >= 2 * d:                                    # P of a span that already reaches bit 0 is never used again
                P2[i] = nl.gate(AND, P[i], P[i - d])
        G, P = G2, P2
        d *= 2
    outs = [p[0]] + [nl.gate(XOR, p[i], G[i - 1]) for i in range(1, nbits)]   # carry into bit i = G[i-1] (span i-1..0)
    outs.append(G[nbits - 1])
    nl.outputs = outs
    return nl


def side_by_side(netlist, k):
    """k disjoint copies of a netlist in one: inputs of copy c are inputs [c * n_inputs, (c + 1) * n_inputs), outputs are
    concatenated in copy order.  Levels become k times as wide (a SIMD batch of the same circuit on independent operands) --
    the shape on which sharding a level over the GPUs of a group pays."""
    nl = Netlist()
    nl.add_inputs(k * netlist.n_inputs)
    inner = netlist.n_wires - netlist.n_inputs
    base = k * netlist.n_inputs
    nl.n_wires += k * inner
    for c in range(k):
        m = lambda w, c=c: c * netlist.n_inputs + w if w < netlist.n_inputs else base + c * inner + (w - netlist.n_inputs)
        for w, bit in netlist.consts.items():
            nl.consts[m(w)] = bit
        for op, a, b, out in netlist.gates:
            nl.gates.append((op, m(a), m(b) if b >= 0 else -1, m(out)))
        nl.outputs += [m(w) for w in netlist.outputs]
    return nl


def expr_to_netlist(expr):
    """Compile a LogicExpr with the TFHE Logip mapping (native and/or/xor/not gates, nander/src/lib.rs:40-62)."""
    nl = Netlist()

    def rec(e):
        if e.kind == "leaf":
            return nl.const(e.value)
        a = rec(e.lhs)
        if e.kind == "not":
            return nl.gate(NOT, a)
        b = rec(e.rhs)
        return nl.gate({"nand": NAND, "and": AND, "or": OR, "xor": XOR}[e.kind], a, b)

    nl.outputs = [rec(expr)]
    return nl


# ------------------------------------------------------------------------------------------------------------
# batched evaluation
# ------------------------------------------------------------------------------------------------------------
def evaluate(engine, netlist, inputs=None, stats=None):
    """Level-synchronous evaluation on one device context.  `inputs`: uint32 [n_inputs][n+1] ciphertexts.
    Host-side wire table + ONE engine call per level (tfhe_b200_gate_batch_mixed when the level mixes opcodes).
    Returns the output ciphertexts."""
    W = K.n + 1
    wires = np.zeros((netlist.n_wires, W), np.uint32)
    if netlist.n_inputs:
        wires[:netlist.n_inputs] = np.ascontiguousarray(inputs, np.uint32).reshape(netlist.n_inputs, W)
    for w, bit in netlist.consts.items():
        wires[w, 0] = 0x20000000 if bit else 0xE0000000
    levels = netlist.levels()
    hist = []
    mixed = getattr(engine, "gate_batch_mixed", None)
    for lev in levels:
        width = sum(len(o) for (_, _, o) in lev.values())
        if mixed is not None and len(lev) > 1:   # gates of different kinds: still ONE launch for the level
            ops = np.concatenate([np.full(len(o), op, np.uint8) for op, (_, _, o) in lev.items()])
            i0 = np.concatenate([a for (a, _, _) in lev.values()])
            i1 = np.concatenate([b for (_, b, _) in lev.values()])
            o = np.concatenate([c for (_, _, c) in lev.values()])
            wires[o] = mixed(ops, wires[i0], wires[i1])
        else:
            for op, (i0, i1, o) in lev.items():
                wires[o] = engine.gate_batch(op, wires[i0], None if op == NOT else wires[i1])
        hist.append(width)
    if stats is not None:
        stats["levels"] = len(levels)
        stats["width_histogram"] = hist
        stats["gates"] = len(netlist.gates)
    return wires[netlist.outputs] if netlist.outputs else wires


class DeviceCircuit:
    """A levelised netlist resident on the device (tfhe_b200_circuit_*): `run(inputs)` uploads the input ciphertexts into the
    wire table, enqueues one launch pair per level with no host round trip in between, and downloads the outputs."""

    def __init__(self, engine, netlist):
        import ctypes as C
        self.engine, self.netlist = engine, netlist
        sizes, ops, i0, i1, o = _flatten_levels(netlist)
        self._sizes = (C.c_size_t * max(1, len(sizes)))(*sizes)
        self._h = C.c_void_p()
        self.levels, self.gates, self.sizes = len(sizes), int(sum(sizes)), sizes
        rc = engine._l.tfhe_b200_circuit_create(engine._ctx, len(sizes), self._sizes, K.ptr(ops), K.ptr(i0), K.ptr(i1), K.ptr(o),
                                                netlist.n_wires, C.byref(self._h))
        engine._ck(rc)

    def run(self, inputs=None):
        import ctypes as C
        import torch
        nl, W = self.netlist, K.n + 1
        wires = np.zeros((nl.n_wires, W), np.uint32)
        if nl.n_inputs:
            wires[:nl.n_inputs] = np.ascontiguousarray(inputs, np.uint32).reshape(nl.n_inputs, W)
        for w, bit in nl.consts.items():
            wires[w, 0] = 0x20000000 if bit else 0xE0000000
        dev = torch.device("cuda", self.engine.device)
        d = torch.from_numpy(wires.view(np.int32)).to(dev)
        st = torch.cuda.current_stream(dev)
        self.engine._ck(self.engine._l.tfhe_b200_circuit_run_device(self.engine._ctx, self._h, C.c_void_p(d.data_ptr()), C.c_void_p(st.cuda_stream)))
        out = d[torch.as_tensor(nl.outputs, device=dev)] if nl.outputs else d
        return out.cpu().numpy().view(np.uint32)

    def close(self):
        if self._h:
            self.engine._l.tfhe_b200_circuit_destroy(self.engine._ctx, self._h)
            self._h = None


def _flatten_levels(netlist):
    """The levels of a netlist as the concatenated arrays the C ABI takes: sizes, ops, in0, in1, out."""
    sizes, ops, i0, i1, o = [], [], [], [], []
    for lev in netlist.levels():
        n = 0
        for op, (a, b, c) in lev.items():
            ops.append(np.full(len(c), op, np.uint8)); i0.append(a); i1.append(b); o.append(c); n += len(c)
        sizes.append(n)
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dt)
    return sizes, cat(ops, np.uint8), cat(i0, np.int32), cat(i1, np.int32), cat(o, np.int32)


def level_plan(sizes, world, shard_min):
    """What tfhe_b200_group_circuit_run does with each level on `world` devices: ("replicated", width) when the level is
    evaluated by every device on its own wire table, or ("sharded", [(first, count) per device]) when it is cut into
    contiguous shards whose outputs are exchanged (the first width % world devices take one gate more)."""
    plan = []
    for w in sizes:
        if world == 1 or w < shard_min:
            plan.append(("replicated", w))
        else:
            base, rem = divmod(w, world)
            plan.append(("sharded", [(r * base + min(r, rem), base + (1 if r < rem else 0)) for r in range(world)]))
    return plan


class GroupCircuit:
    """A levelised netlist on every GPU of a DeviceGroup (tfhe_b200_group_circuit_*): wide levels are sharded over the devices
    and their outputs exchanged over NCCL, narrow ones are evaluated by every device.  `run(inputs)` returns the
    output ciphertexts; `last` holds how many levels of the run were sharded / replicated."""

    def __init__(self, group, netlist, shard_min=0):
        import ctypes as C
        self.group, self.netlist = group, netlist
        sizes, ops, i0, i1, o = _flatten_levels(netlist)
        self._sizes = (C.c_size_t * max(1, len(sizes)))(*sizes)
        self._h = C.c_void_p()
        self.levels, self.gates, self.sizes = len(sizes), int(sum(sizes)), sizes
        group._ck(group._l.tfhe_b200_group_circuit_create(group._g, len(sizes), self._sizes, K.ptr(ops), K.ptr(i0), K.ptr(i1), K.ptr(o),
                                                         netlist.n_wires, shard_min, C.byref(self._h)))
        self.last = {}

    def run(self, inputs=None):
        import ctypes as C
        nl, W = self.netlist, K.n + 1
        ins = np.ascontiguousarray(inputs, np.uint32).reshape(nl.n_inputs, W) if nl.n_inputs else np.zeros((0, W), np.uint32)
        cw = np.ascontiguousarray(list(nl.consts.keys()), np.int32)
        cb = np.ascontiguousarray(list(nl.consts.values()), np.uint8)
        ow = np.ascontiguousarray(nl.outputs if nl.outputs else np.arange(nl.n_wires), np.int32)
        out = np.empty((len(ow), W), np.uint32)
        g = self.group
        g._ck(g._l.tfhe_b200_group_circuit_run(g._g, self._h, K.ptr(ins), len(ins), K.ptr(cw), K.ptr(cb), len(cw), K.ptr(ow), len(ow), K.ptr(out)))
        sh, rep, smin = C.c_uint64(), C.c_uint64(), C.c_size_t()
        g._l.tfhe_b200_group_circuit_stats(self._h, C.byref(sh), C.byref(rep), C.byref(smin))
        self.last = {"sharded_levels": sh.value, "replicated_levels": rep.value, "shard_min": smin.value}
        return out

    def close(self):
        if self._h:
            self.group._l.tfhe_b200_group_circuit_destroy(self.group._g, self._h)
            self._h = None


def eval_logic_expr(pros, expr):
    """eval_logic_expr(&pros, exp) (nander/src/lib.rs:72-89) for pros = TFHE, evaluated level by level in batches."""
    return evaluate(pros.engine, expr_to_netlist(expr))
