"""Host-checkable restatement of the division-free pipeline bookkeeping of the tcgen05 kernels (round 2): the CUDA code keeps
ring positions and the work-item decomposition as incrementally updated integers (csrc/umma.cuh: RingPos; csrc/umma_tap.cuh:
TapIter and the MMA issuer's window / skip / release logic).  These tests re-implement those update rules line by line in
Python and check them against the closed forms they replaced (`g % S`, `(g / S) & 1`, `item % n_tsplit`, ... and the
slice-window definition of the tap GEMM), over randomised shapes including the data-gradient launches (t0 < 0) and the
output-time split.  The GPU parity tests exercise the kernels themselves; this pins the arithmetic on a CPU box."""
import random


class RingPos:                      # csrc/umma.cuh: struct RingPos
    def __init__(self, s=0, ph=0):
        self.s, self.ph = s, ph

    def advance(self, S):
        self.s += 1
        if self.s == S:
            self.s, self.ph = 0, self.ph ^ 1

    def advance_by(self, n, S):
        self.s += n
        while self.s >= S:
            self.s -= S
            self.ph ^= 1

    def copy(self):
        return RingPos(self.s, self.ph)


def closed(g, S):
    return g % S, (g // S) & 1


def test_ring_position_matches_division():
    rnd = random.Random(0)
    for _ in range(200):
        S = rnd.randint(1, 12)
        rp, g = RingPos(), 0
        for _ in range(300):
            assert (rp.s, rp.ph) == closed(g, S)
            if rnd.random() < 0.5:
                rp.advance(S); g += 1
            else:
                n = rnd.randint(0, 3 * S)
                rp.advance_by(n, S); g += n


def test_item_iterator_matches_division():
    """csrc/umma_tap.cuh: TapIter::next with the host-side stride decomposition (launch_tap: d_ts, d_nt, d_b)."""
    rnd = random.Random(1)
    for _ in range(300):
        n_tsplit, nnt, B, G = rnd.randint(1, 4), rnd.randint(1, 17), rnd.randint(1, 300), rnd.randint(1, 148)
        n_items = B * nnt * n_tsplit
        d_ts, d_nt, d_b = G % n_tsplit, (G // n_tsplit) % nnt, G // (n_tsplit * nnt)
        for blk in rnd.sample(range(G), min(G, 5)):
            item = blk
            ts, rest = item % n_tsplit, item // n_tsplit
            b, nt = rest // nnt, rest % nnt                       # the kernel-start decomposition (shuffled once)
            while item < n_items:
                rest = item // n_tsplit
                assert (ts, nt, b) == (item % n_tsplit, rest % nnt, rest // nnt)
                item += G                                         # TapIter::next
                ts += d_ts
                c = 0
                if ts >= n_tsplit:
                    ts -= n_tsplit; c = 1
                nt += d_nt + c
                c = 0
                if nt >= nnt:
                    nt -= nnt; c = 1
                b += d_b + c


def _items(T_out, T_src, Kt, t0, n_tsplit):
    chunk = (T_out + n_tsplit - 1) // n_tsplit
    out = []
    for ts in range(n_tsplit):
        t_begin = ts * chunk
        t_end = min(t_begin + chunk, T_out)
        if t_begin >= t_end:
            continue
        s_lo = max(t_begin + t0, 0)
        s_hi = min(t_end + t0 + Kt - 1, T_src)
        out.append((t_begin, t_end, s_lo, s_hi))
    return out


def test_issuer_window_bookkeeping_matches_slice_definition():
    """The MMA issuer of umma_tap_kernel: per output step it must (a) multiply exactly the slices ti = t_o + j + t0 that lie
    in [0, T_src), each found at ring position of global slice counter g_base + (ti - s_lo); (b) wait for a slice's full
    barrier before its first use and never again; (c) release every slice exactly once, after its last use."""
    rnd = random.Random(2)
    for _ in range(400):
        Kt = rnd.randint(1, 4)
        T_src = rnd.randint(Kt, 12)
        if rnd.random() < 0.5:
            t0, T_out = 0, T_src - Kt + 1                         # forward conv / linear map
        else:
            t0, T_out = -(Kt - 1), T_src + Kt - 1                 # data gradient
        S = rnd.randint(Kt + 1, 12)
        n_tsplit = rnd.randint(1, min(4, T_out))
        base, g_base = RingPos(), 0
        for rep in range(3):                                      # several items in a row (ring wraps across items)
            for (t_begin, t_end, s_lo, s_hi) in _items(T_out, T_src, Kt, t0, n_tsplit):
                win = base.copy()
                skip = s_lo - (t_begin + t0)
                n_waited = 0
                waited, released, last_use = set(), [], {}
                for t_o in range(t_begin, t_end):
                    pos = win.copy()
                    d_win = 0 if skip > 0 else t_o + t0 - s_lo
                    used = []
                    j = max(skip, 0)
                    while j < Kt:
                        ti = t_o + j + t0
                        if ti >= s_hi:
                            break
                        d = d_win + j - max(skip, 0)
                        if d >= n_waited:
                            assert d == n_waited, "slices are waited for in order"
                            waited.add(d); n_waited = d + 1
                        assert d in waited
                        assert d == ti - s_lo and (pos.s, pos.ph) == closed(g_base + d, S)
                        used.append(ti); last_use[ti] = t_o
                        j += 1; pos.advance(S)
                    assert used == [t_o + j + t0 for j in range(Kt) if 0 <= t_o + j + t0 < T_src]
                    if t_o == t_end - 1:
                        r, ti = win.copy(), s_lo + d_win
                        while ti < s_hi:
                            assert (r.s, r.ph) == closed(g_base + ti - s_lo, S)
                            released.append(ti); ti += 1; r.advance(S)
                    elif skip <= 0 and t_o + t0 < s_hi:
                        assert (win.s, win.ph) == closed(g_base + t_o + t0 - s_lo, S)
                        released.append(t_o + t0)
                    for ti in released:                            # never released before its last use
                        assert last_use.get(ti, -1) <= t_o
                    if skip > 0:
                        skip -= 1
                    else:
                        win.advance(S)
                assert sorted(released) == list(range(s_lo, s_hi)) and len(set(released)) == len(released)
                assert waited == set(range(s_hi - s_lo))
                base.advance_by(s_hi - s_lo, S); g_base += s_hi - s_lo
                assert (base.s, base.ph) == closed(g_base, S)


def test_fb2_accumulator_ring_has_one_writer_and_one_reader_at_a_time():
    """umma_fb2_kernel: data-gradient accumulators X_tau (ring of 6).  Tile t writes X_t, X_t+1, X_t+2 (X_t+2 fresh; all
    three fresh at t = 0), commits X_t complete (the last tile also X_t+1, X_t+2); E2 drains them in order.  Simulates the
    issuer / E2 counters over several items and checks that a fresh write never lands on an undrained accumulator and
    that every accumulator is complete exactly when E2 reads it."""
    NX = 6
    for T2 in range(1, 9):
        T1 = T2 + 2
        xw, drained, state = RingPos(), 0, {}
        written = 0                                               # global input-step counter of the next fresh accumulator
        for item in range(4):
            contrib = {tau: 0 for tau in range(T1)}
            for t in range(T2):
                xp = xw.copy()
                for j in range(3):
                    tau = t + j
                    fresh = j == 2 or t == 0
                    g = item * T1 + tau
                    assert (xp.s, xp.ph) == closed(g, NX)
                    if fresh:
                        assert g - drained < NX, "would overwrite an accumulator E2 has not drained"
                        state[xp.s] = g
                    assert state[xp.s] == g
                    contrib[tau] += 1
                    xp.advance(NX)
                done = [t] if t < T2 - 1 else [t, t + 1, t + 2]
                for tau in done:
                    expect = sum(1 for j in range(3) if 0 <= tau - j < T2)
                    assert contrib[tau] == expect
                    assert item * T1 + tau == drained             # E2 drains in commit order
                    drained += 1
                xw.advance(NX)
            xw.advance_by(2, NX)
            written += T1
            assert (xw.s, xw.ph) == closed(written, NX) and drained == written
