"""Random small CSR NFAs in the reference's BRAM image format (test infrastructure)."""
import numpy as np


def build_entries(rows):
    """rows[s] = list of (symbol, target).  Returns the BRAM image: row_ptr | transitions | zero pad
    (Design/FPGA.v:773,782,793,888-898)."""
    n = len(rows)
    rp = [0]
    tr = []
    for r in rows:
        seen = sorted(set(r), key=lambda st: (st[1], st[0]))   # rows sorted by target, no duplicates
        for sym, tgt in seen:
            tr.append(((sym & 0xFF) << 24) | (tgt & 0xFFFFFF))
        rp.append(len(tr))
    e = rp + tr
    while len(e) % 4:
        e.append(0)
    return np.array(e, dtype=np.uint32), n


def random_nfa(rng, n_states=40, alphabet=8, p_accept=0.15, p_sticky=0.1, p_branch=0.3, max_fanout=3,
               wide_classes=True, unanchored=False):
    """Mix of the structures the shipped rulesets have: chains on symbol pairs, branching tries, class
    edges, self-loops on (nearly) all symbols, zero-out-degree accept states, same-symbol multi-target."""
    syms = rng.choice(256, size=alphabet, replace=False)
    rows = []
    for s in range(n_states):
        r = []
        kind = rng.random()
        if s > 1 and kind < p_accept:
            rows.append(r)
            continue
        if kind < p_accept + p_sticky:  # sticky: self-loop on all (or all but a few) symbols
            drop = set(rng.choice(256, size=rng.integers(0, 3), replace=False).tolist())
            r += [(c, s) for c in range(256) if c not in drop]
            for _ in range(rng.integers(1, 4)):
                r.append((int(rng.choice(syms)), int(rng.integers(1, n_states))))
        elif kind < p_accept + p_sticky + p_branch:  # branching
            for _ in range(rng.integers(2, 7)):
                c = int(rng.choice(syms))
                for _ in range(rng.integers(1, max_fanout + 1)):
                    r.append((c, int(rng.integers(1, n_states))))
            if rng.random() < 0.3:  # small self loop (\s+ style)
                for c in rng.choice(syms, size=2):
                    r.append((int(c), s))
        else:  # chain: one or two symbols (case pair) -> one target, sometimes a wide class
            t = int(rng.integers(1, n_states))
            if wide_classes and rng.random() < 0.2:
                lo = int(rng.integers(0, 200))
                r += [(c, t) for c in range(lo, lo + int(rng.integers(3, 56)))]
            else:
                c = int(rng.choice(syms))
                r.append((c, t))
                if rng.random() < 0.6:
                    r.append((c ^ 0x20, t))
        rows.append(r)
    # state 0: start; make sure something is reachable
    if not rows[0]:
        rows[0] = [(int(rng.choice(syms)), 1)]
    if unanchored and n_states >= 3:
        # the shape of the shipped rulesets: state 0 enters a ".*" state on every symbol; that state stays active
        # for ever and starts every pattern (the library follows its successors with a start DFA)
        rows[0] = [(c, 1) for c in range(256)]
        rows[1] = [(c, 1) for c in range(256)]
        for _ in range(int(rng.integers(2, 10))):
            c = int(rng.choice(syms))
            for _ in range(int(rng.integers(1, max_fanout + 1))):
                rows[1].append((c, int(rng.integers(2, n_states))))
    return build_entries(rows), syms


def random_streams(rng, syms, n_streams, length, p_alpha=0.85):
    """Bytes drawn mostly from the NFA's alphabet so that sets stay active."""
    pick = rng.random((n_streams, length)) < p_alpha
    a = rng.choice(syms, size=(n_streams, length))
    b = rng.integers(0, 256, size=(n_streams, length))
    return np.where(pick, a, b).astype(np.uint8)
