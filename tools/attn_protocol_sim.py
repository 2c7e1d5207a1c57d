"""Discrete-event model of the mbarrier protocol of the persistent attention kernel (csrc/attn_tc.cu): TMA warp, MMA
warp and the eight softmax warps as coroutines over mbarriers with hardware phase-parity semantics.  Checks, for many
item sequences (two-tile / single-tile items, 1..n key tiles, warps without rows), that
  * nothing deadlocks,
  * every wait is released by the completion it MEANS (no parity aliasing: a wait for completion c never passes on
    completion c-2 and is never left behind by completion c+2),
  * TMEM / shared-memory buffers are never overwritten before their readers are done (S, P, O, Q halves, K/V stages).
Runs on the CPU in seconds: `python tools/attn_protocol_sim.py`.  Transcribed from the kernel's role code — keep in sync."""
import itertools
import random
import sys

K_STAGES = V_STAGES = 4


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: too many arrivals in a phase"
        if self.pending == 0:
            self.completed += 1
            self.pending = self.count

    def passes(self, parity):
        # hardware: try_wait.parity(p) succeeds iff the phase with parity p has completed = current phase parity != p
        return (self.completed & 1) != parity


class Wait:
    def __init__(self, bar, index):
        self.bar, self.index = bar, index  # index = the completion (0-based) the waiter means; -1 = "fresh pass"


def simulate(items, n_kv, rows_of, seed=0, verbose=False):
    """items: list of bool (two-tile?) for ONE CTA in order.  rows_of(item, t, sub) -> bool (warp has rows)."""
    rnd = random.Random(seed)
    B = {}
    for n, c in (("q_full", 1), ("q_empty", 1)):
        for i in range(2):
            B[f"{n}{i}"] = Bar(f"{n}{i}", c)
    for s in range(K_STAGES):
        B[f"k_full{s}"], B[f"k_empty{s}"] = Bar(f"k_full{s}", 1), Bar(f"k_empty{s}", 1)
    for s in range(V_STAGES):
        B[f"v_full{s}"], B[f"v_empty{s}"] = Bar(f"v_full{s}", 1), Bar(f"v_empty{s}", 1)
    for t in range(2):
        B[f"s_full{t}"], B[f"s_empty{t}"] = Bar(f"s_full{t}", 1), Bar(f"s_empty{t}", 4)
        B[f"p_full{t}"], B[f"pv_done{t}"] = Bar(f"p_full{t}", 4), Bar(f"pv_done{t}", 1)
    for q in range(4):
        B[f"turn_a{q}"], B[f"turn_b{q}"] = Bar(f"turn_a{q}", 1), Bar(f"turn_b{q}", 1)

    # resource state for hazard checks
    state = {"S": [None, None], "P": [None, None], "Qhalf": [None, None], "Kst": [None] * K_STAGES,
             "Vst": [None] * V_STAGES}
    async_q = []  # (due_time, seq, fn): MMA / TMA completions, in order per engine
    now = [0]
    ctr = itertools.count()

    def later(delay, fn):
        async_q.append((now[0] + delay, next(ctr), fn))

    def wait(bar, index):
        return Wait(B[bar], index)

    # ------------------------------------------------------------------ roles (generators yield Wait objects)
    def tma():
        kt = 0
        for it, two in enumerate(items):
            qb = it & 1
            yield wait(f"q_empty{qb}", (it >> 1) - 1)
            assert state["Qhalf"][qb] in (None, "free"), f"Q half {qb} overwritten while in use"
            state["Qhalf"][qb] = "loading"
            later(rnd.randint(3, 30), lambda qb=qb, it=it: (state["Qhalf"].__setitem__(qb, it), B[f"q_full{qb}"].arrive()))
            for j in range(n_kv):
                sk, sv = kt % K_STAGES, kt % V_STAGES
                yield wait(f"k_empty{sk}", kt // K_STAGES - 1)
                later(rnd.randint(3, 30), lambda sk=sk, kt=kt: (state["Kst"].__setitem__(sk, kt), B[f"k_full{sk}"].arrive()))
                yield wait(f"v_empty{sv}", kt // V_STAGES - 1)
                later(rnd.randint(3, 30), lambda sv=sv, kt=kt: (state["Vst"].__setitem__(sv, kt), B[f"v_full{sv}"].arrive()))
                kt += 1

    mma_chain = [0]  # completion time of the last MMA batch (tensor pipe is in order)

    def mma():
        n_qk, n_pv = [0, 0], [0, 0]

        def commit(delay, fns):
            t = max(mma_chain[0], now[0]) + delay
            mma_chain[0] = t
            async_q.append((t, next(ctr), lambda: [f() for f in fns]))

        def issue_qk(qb, kt_idx, j, t, item_seq, release_k, release_q):
            if n_qk[t] > 0:
                yield wait(f"s_empty{t}", n_qk[t] - 1)
            assert state["Qhalf"][qb] == item_seq, f"QK reads Q half {qb}: holds {state['Qhalf'][qb]}, wants item {item_seq}"
            assert state["Kst"][kt_idx % K_STAGES] == kt_idx, "QK reads a K stage that does not hold its tile"
            tag = (item_seq, j)
            fns = [lambda: state["S"].__setitem__(t, tag), lambda: B[f"s_full{t}"].arrive()]
            if release_k:
                fns.append(lambda: B[f"k_empty{kt_idx % K_STAGES}"].arrive())
            if release_q:
                fns.append(lambda: (state["Qhalf"].__setitem__(qb, "free"), B[f"q_empty{qb}"].arrive()))
            commit(rnd.randint(5, 20), fns)
            n_qk[t] += 1

        def issue_pv(kt_idx, j, t, item_seq, release_v):
            yield wait(f"p_full{t}", n_pv[t])
            assert state["P"][t] == (item_seq, j), f"PV({item_seq},{j}) tile {t} reads P = {state['P'][t]}"
            assert state["Vst"][kt_idx % V_STAGES] == kt_idx, "PV reads a V stage that does not hold its tile"
            fns = [lambda: B[f"pv_done{t}"].arrive()]
            if release_v:
                fns.append(lambda: B[f"v_empty{kt_idx % V_STAGES}"].arrive())
            commit(rnd.randint(5, 20), fns)
            n_pv[t] += 1

        if not items:
            return
        kt, it = 0, 0
        yield wait("q_full0", 0)
        yield wait("k_full0", 0)
        yield from issue_qk(0, 0, 0, 0, 0, not items[0], (not items[0]) and n_kv == 1)
        if items[0]:
            yield from issue_qk(0, 0, 0, 1, 0, True, n_kv == 1)
        while True:
            two = items[it]
            has_next = it + 1 < len(items)
            qb = it & 1
            for j in range(n_kv):
                in_item = j + 1 < n_kv
                follow = in_item or has_next
                f_two = two if in_item else (items[it + 1] if has_next else False)
                fqb = qb if in_item else qb ^ 1
                fj = j + 1 if in_item else 0
                f_seq = it if in_item else it + 1
                f_last = fj == n_kv - 1
                if follow:
                    if not in_item:
                        yield wait(f"q_full{fqb}", (it + 1) >> 1)
                    yield wait(f"k_full{(kt + 1) % K_STAGES}", (kt + 1) // K_STAGES)
                    yield from issue_qk(fqb, kt + 1, fj, 0, f_seq, not f_two, (not f_two) and f_last)
                if two and j >= 1:
                    yield from issue_pv(kt - 1, j - 1, 1, it, True)
                if follow and f_two:
                    yield from issue_qk(fqb, kt + 1, fj, 1, f_seq, True, f_last)
                yield wait(f"v_full{kt % V_STAGES}", kt // V_STAGES)
                yield from issue_pv(kt, j, 0, it, not two)
                kt += 1
            if two:
                yield from issue_pv(kt - 1, n_kv - 1, 1, it, True)
            if not has_next:
                break
            it += 1

    def softmax(t, sub):
        n_t, n_tok = 0, 0
        for seq, two in enumerate(items):
            if t == 1 and not two:
                continue
            nb = n_t
            has_rows = rows_of(seq, t, sub)

            def take_turn():
                if two:
                    if t == 0:
                        if n_tok > 0:
                            yield wait(f"turn_b{sub}", n_tok - 1)
                    else:
                        yield wait(f"turn_a{sub}", n_tok)

            if not has_rows:
                for j in range(n_kv):
                    yield wait(f"s_full{t}", nb + j)
                    B[f"s_empty{t}"].arrive()
                    yield from take_turn()
                    if two:
                        B[f"turn_{'a' if t == 0 else 'b'}{sub}"].arrive()
                        n_tok += 1
                    if nb + j > 0:
                        yield wait(f"pv_done{t}", nb + j - 1)
                    B[f"p_full{t}"].arrive()
                yield wait(f"pv_done{t}", nb + n_kv - 1)
                n_t = nb + n_kv
                continue
            yield wait(f"s_full{t}", nb)
            for j in range(n_kv):
                assert state["S"][t] == (seq, j), f"softmax t{t} sub{sub} item {seq} tile {j} reads S = {state['S'][t]}"
                B[f"s_empty{t}"].arrive()
                yield from take_turn()
                yield ("work", rnd.randint(1, 10))
                if two:
                    B[f"turn_{'a' if t == 0 else 'b'}{sub}"].arrive()
                    n_tok += 1
                yield ("work", rnd.randint(1, 10))
                if j + 1 < n_kv:
                    yield wait(f"s_full{t}", nb + j + 1)
                elif j > 0:
                    yield wait(f"pv_done{t}", nb + j - 1)
                # P store (all four warps store their lanes; tag once per phase for the hazard check)
                state["P"][t] = (seq, j)
                B[f"p_full{t}"].arrive()
            yield wait(f"pv_done{t}", nb + n_kv - 1)
            yield ("work", rnd.randint(1, 5))  # epilogue reads O
            n_t = nb + n_kv

    roles = {"tma": tma(), "mma": mma()}
    for t in range(2):
        for sub in range(4):
            roles[f"sm{t}{sub}"] = softmax(t, sub)
    blocked = {}      # role -> Wait
    sleeping = {}     # role -> wake time
    done = set()
    steps = 0
    while len(done) < len(roles):
        steps += 1
        assert steps < 2_000_000, "runaway"
        progressed = False
        order = list(roles)
        rnd.shuffle(order)
        for name in order:
            if name in done:
                continue
            if name in sleeping:
                if sleeping[name] > now[0]:
                    continue
                del sleeping[name]
            gen = roles[name]
            w = blocked.get(name)
            if w is not None:
                if w.index < 0 or w.bar.passes(w.index & 1):
                    if w.index >= 0:
                        assert w.bar.completed == w.index + 1, (f"{name}: wait on {w.bar.name} for completion {w.index} "
                                                                f"released with {w.bar.completed} completed (parity aliasing)")
                    del blocked[name]
                else:
                    assert w.bar.completed <= w.index, (f"{name}: wait on {w.bar.name} for completion {w.index} left behind "
                                                        f"({w.bar.completed} completed)")
                    continue
            try:
                ev = next(gen)
            except StopIteration:
                done.add(name)
                progressed = True
                continue
            progressed = True
            if isinstance(ev, Wait):
                blocked[name] = ev
            else:
                sleeping[name] = now[0] + ev[1]
        # fire due asynchronous completions (in time order)
        async_q.sort()
        while async_q and async_q[0][0] <= now[0]:
            _, _, fn = async_q.pop(0)
            fn()
            progressed = True
        if not progressed or all((n in blocked or n in sleeping or n in done) for n in roles):
            nxt = [t for t, _, _ in async_q] + list(sleeping.values())
            if not nxt:
                still = {n: (blocked[n].bar.name, blocked[n].index, blocked[n].bar.completed) for n in blocked}
                if all(not (blocked[n].index < 0 or blocked[n].bar.passes(blocked[n].index & 1)) for n in blocked) and len(done) < len(roles):
                    raise AssertionError(f"DEADLOCK: {still}")
            else:
                now[0] = max(now[0] + 1, min(nxt))
    return steps


def main():
    n = 0
    for n_kv in (1, 2, 3, 5, 9):
        for items in itertools.chain(itertools.product([True, False], repeat=1), itertools.product([True, False], repeat=2),
                                     itertools.product([True, False], repeat=3), [(True,) * 6, (True,) * 5 + (False,),
                                     (True, True, False, False), (False,) * 4, (True, False, True, False, True)]):
            for rows_mode in ("all", "first_warp_only", "tile1_partial"):
                def rows_of(seq, t, sub, items=items, rows_mode=rows_mode):
                    if rows_mode == "all":
                        return True
                    if rows_mode == "first_warp_only":   # ragged single tile with one valid row
                        return sub == 0 if not items[seq] else True
                    return not (t == 1 and sub >= 1 and seq == len(items) - 1)   # last item's tile 1 has 17 rows
                for seed in range(3):
                    simulate(list(items), n_kv, rows_of, seed=seed)
                    n += 1
    print(f"attention barrier protocol: {n} simulated schedules, no deadlock, no parity aliasing, no buffer hazard")


if __name__ == "__main__":
    sys.exit(main())
