"""Random well-formed tables (SPEC.md section 7): the table INTERPRETER kernels against Oracle B.

The shipped games exercise a handful of (kind, action, exit, entry, branch) combinations; the C ABI accepts any
well-formed table.  Tables are drawn directly in the binary format (not through the DSL compiler) from a seeded
generator, so failures reproduce."""
import numpy as np
import pytest

from game_engine_b200 import table as T


def _pred(rng, n_fields):
    """A random two-clause DNF over the family's mask fields (SPEC section 2)."""
    def clause():
        pos = neg = 0
        for f in rng.choice(n_fields, size=rng.integers(0, 3), replace=False):
            if rng.random() < 0.6:
                pos |= 1 << int(f)
            else:
                neg |= 1 << int(f)
        return pos, neg
    c0 = clause()
    c1 = clause() if rng.random() < 0.35 else T.CLAUSE_EMPTY
    return c0 + c1


def random_table(seed: int, family: int) -> T.Table:
    rng = np.random.default_rng(seed)
    wolf = family == T.FAMILY_WEREWOLF
    P = int(rng.integers(4, 33)) if wolf else int(rng.integers(2, 33))
    if rng.random() < 0.5:
        P = int(rng.choice([4, 8, 16, 32] if not wolf else [5, 8, 16, 32]))
    n_ph = int(rng.integers(4, 15))
    fields = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12] if wolf else [0, 1, 2, 3, 4]
    tab = T.Table(family=family, n_players=P, n_wolves=int(rng.integers(1, max(2, P // 3))) if wolf else 0,
                  rounds=0 if wolf else int(rng.integers(1, 3)), max_revotes=int(rng.integers(0, 3)) if wolf else 0,
                  init_masks=(0b11 | (int(rng.integers(0, 2)) << 2)) if wolf else (1 << 3 | int(rng.integers(0, 2))))
    # comparison fields (numeric conditions) and the mask-field ids they occupy
    n_cmp = int(rng.integers(0, T.MAX_CMP[family] + 1)) if rng.random() < 0.6 else 0
    tab.cmps = [(0 if wolf else int(rng.integers(0, 3)), int(rng.integers(0, 6)), int(rng.integers(0, 4 if not wolf else P + 1)))
                for _ in range(n_cmp)]
    fields = fields + [T.cmp_field_id(family, k) for k in range(n_cmp)]
    tab.preds = [_pred(rng, fields) for _ in range(int(rng.integers(3, 9)))]
    for i in range(len(tab.preds) - 1):                             # some predicates continue in the next record
        if rng.random() < 0.25:
            p0 = tab.preds[i]
            tab.preds[i] = (p0[0] | T.PRED_CONTINUED,) + tuple(p0[1:])
    tab.preds.append((1, 0) + T.CLAUSE_EMPTY)                       # "field 0" (alive / speaker): keeps games lively
    npred = len(tab.preds)
    exits = [T.EX_VOTE_KILL, T.EX_PROTECT, T.EX_INVESTIGATE_RESOLVE, T.EX_DAY_VOTE] if wolf else \
        [T.EX_T_STATEMENTS, T.EX_T_LIE, T.EX_T_VOTES]
    entries = [T.EN_ASSIGN_ROLES, T.EN_NIGHT_RESET] if wolf else [T.EN_T_ROUND_START, T.EN_T_REVEAL, T.EN_T_SCORE, T.EN_T_FINAL]
    for i in range(n_ph):
        last = i == n_ph - 1
        kind = T.KIND_TERMINAL if last else int(rng.choice([T.KIND_UI, T.KIND_TIMER, T.KIND_ACTION, T.KIND_ACTION]))
        ph = T.Phase(id=i if not last else 99, kind=kind)
        if rng.random() < (0.45 if i else 0.0):
            ph.entry_op = int(rng.choice(entries))
        if i == 1 and wolf:
            ph.entry_op = T.EN_ASSIGN_ROLES                          # most werewolf predicates need roles
        if kind == T.KIND_ACTION:
            ph.actor_pred = int(rng.integers(0, npred))
            ph.action_op = int(rng.choice([T.ACT_PICK_PLAYER, T.ACT_PICK_OPTION, T.ACT_MARK]))
            if rng.random() < 0.75:
                ph.exit_op = int(rng.choice(exits))
                if wolf:
                    ph.action_op = T.ACT_PICK_PLAYER
            if ph.action_op == T.ACT_PICK_PLAYER:
                ph.action_arg = int(rng.integers(0, npred))
                ph.action_flags = int(rng.integers(0, 2))
            elif ph.action_op == T.ACT_PICK_OPTION:
                ph.action_arg = int(rng.integers(1, 7))
        if kind != T.KIND_TERMINAL:
            nb = int(rng.choice([1, 1, 2, 3, 4]))
            ops = [T.BR_COUNT_EQ0, T.BR_COUNT_GE, T.BR_PREV_IN] + ([T.BR_TIE_PENDING] if wolf else [T.BR_ALL_VAL_GE])
            for b in range(nb):
                op = T.BR_ALWAYS if b == nb - 1 and rng.random() < 0.7 else int(rng.choice(ops + [T.BR_ALWAYS]))
                nxt = int(rng.integers(0, n_ph)) if rng.random() < 0.8 else min(i + 1, n_ph - 1)
                br = T.Branch(op=op, next=nxt, tag=int(rng.integers(0, 3)) if rng.random() < 0.2 else 0)
                if op in (T.BR_COUNT_EQ0, T.BR_COUNT_GE):
                    br.a = int(rng.integers(0, npred))
                if op == T.BR_COUNT_GE:
                    br.arg = int(rng.integers(0, npred))
                if op == T.BR_PREV_IN:
                    br.arg = int(rng.integers(0, 1 << n_ph))
                if op == T.BR_ALL_VAL_GE:
                    br.a, br.arg = int(rng.integers(0, 3)), int(rng.integers(0, 3))
                ph.branches.append(br)
        tab.phases.append(ph)
    return tab


SEEDS = list(range(40))


@pytest.mark.parametrize("family", [T.FAMILY_WEREWOLF, T.FAMILY_TTL])
def test_random_tables_are_accepted_and_deterministic(family):
    """CPU: the generator only draws well-formed tables; Oracle B accepts them, splits and thread counts agree."""
    from oracle.oracle import Oracle
    for seed in SEEDS[:12]:
        tab = random_table(seed, family)
        blob = tab.pack()
        assert T.Table.unpack(blob).pack() == blob
        o = Oracle(blob)
        a, b = o.init(300), o.init(300)
        sa, sb = o.new_stats(), o.new_stats()
        o.step(a, 10, seed, 40, sa, threads=1)
        o.step(b[:111], 10, seed, 40, sb, threads=2)
        o.step(b[111:], 121, seed, 40, sb, threads=3)
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(sa, sb)


@pytest.mark.parametrize("bad", ["exit_on_ui", "foreign_exit", "foreign_entry", "no_action_op", "option_vote", "foreign_branch",
                                 "bad_cmp_op", "undefined_field", "dangling_chain", "too_many_cmps"])
def test_ill_formed_tables_are_rejected_by_the_oracle(bad):
    from oracle.oracle import Oracle
    tab = random_table(3, T.FAMILY_WEREWOLF)
    act = next(p for p in tab.phases if p.kind == T.KIND_ACTION)
    ui = next(p for p in tab.phases if p.kind in (T.KIND_UI, T.KIND_TIMER))
    if bad == "exit_on_ui":
        ui.exit_op = T.EX_DAY_VOTE
    elif bad == "foreign_exit":
        act.exit_op = T.EX_T_VOTES
    elif bad == "foreign_entry":
        ui.entry_op = T.EN_T_SCORE
    elif bad == "no_action_op":
        act.action_op = T.ACT_NONE
    elif bad == "option_vote":
        act.exit_op, act.action_op, act.action_arg = T.EX_DAY_VOTE, T.ACT_PICK_OPTION, 3
    elif bad == "foreign_branch":
        ui.branches[0].op = T.BR_ALL_VAL_GE
    elif bad == "bad_cmp_op":
        tab.cmps = [(0, 9, 1)]
    elif bad == "undefined_field":                      # a predicate on comparison field 14 the table does not define
        tab.cmps = tab.cmps[:1]
        tab.preds[0] = (1 << 14, 0) + T.CLAUSE_EMPTY
    elif bad == "dangling_chain":
        tab.preds[-1] = (tab.preds[-1][0] | T.PRED_CONTINUED,) + tuple(tab.preds[-1][1:])
    blob = tab.pack()
    if bad == "too_many_cmps":                          # the count byte says three; the werewolf family has two
        blob = blob[:13] + bytes([3]) + blob[14:]
    with pytest.raises(ValueError):
        Oracle(blob)


class _Blob:
    """Minimal CompiledGame stand-in for Table(): the C ABI only needs the blob and the record size."""

    def __init__(self, tab):
        self.blob = tab.pack()
        self.record_size = T.record_size(tab.family, tab.n_players)
        self.audience_preds = {}


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["tps", "coop"])
@pytest.mark.parametrize("family", [T.FAMILY_WEREWOLF, T.FAMILY_TTL])
def test_interpreter_kernels_match_oracle_on_random_tables(family, kernel):
    from game_engine_b200.batch import SessionBatch, Table
    from game_engine_b200.capi import GameEngineError
    from oracle.oracle import Oracle
    for seed in SEEDS:
        tab = random_table(seed, family)
        cg = _Blob(tab)
        o = Oracle(cg.blob)
        n, first = 777, (seed << 33) + 5
        b = SessionBatch(Table(cg), n, first_session_id=first, seed=seed + 1, kernel=kernel)
        if kernel == "tps" and seed % 3 == 0:
            b.set_compaction(2, 4)
        if kernel == "tps" and seed % 3 == 1 and family == T.FAMILY_WEREWOLF:
            b.set_regroup(2, 16)
        rec = o.init(n)
        ost = o.new_stats()
        for k in range(48):
            b.step(1)
            o.step(rec, first, seed + 1, 1, ost)
            got = b.export_state()
            if not np.array_equal(got, rec):
                bad = np.nonzero((got != rec).any(axis=1))[0]
                i = int(bad[0])
                raise AssertionError("table seed %d (P=%d, %d phases) step %d: %d/%d sessions differ\n gpu=%s\n cpu=%s\n phases=%s\n preds=%s"
                                     % (seed, tab.n_players, len(tab.phases), k, len(bad), n, got[i].tolist(), rec[i].tolist(),
                                        tab.phases, tab.preds))
        o.stats_final(rec, ost)
        np.testing.assert_array_equal(b.stats(), ost, err_msg="table seed %d" % seed)
        b.close()
    # the library applies the same well-formedness rules as the oracle
    bad = random_table(3, family)
    next(p for p in bad.phases if p.kind in (T.KIND_UI, T.KIND_TIMER)).exit_op = T.EX_DAY_VOTE if family == T.FAMILY_WEREWOLF else T.EX_T_VOTES
    with pytest.raises(GameEngineError):
        Table(_Blob(bad))


def test_library_and_oracle_validators_agree_on_mutated_blobs(games):
    """Random byte mutations of valid tables: ge_table_create (CUDA library, validates on the host) and Oracle B's
    tab_open must take the same accept / reject decision — a blob only one side accepts could not be parity-tested."""
    import ctypes
    from game_engine_b200 import capi
    from oracle.oracle import Oracle
    L = capi.lib()
    rng = np.random.default_rng(99)
    blobs = [games("werewolf-(mafia)", 8).blob, games("two-truths-and-a-lie", 4).blob, games("werewolf-revote", 32).blob,
             random_table(5, T.FAMILY_WEREWOLF).pack(), random_table(6, T.FAMILY_TTL).pack()]
    accepted = rejected = 0
    for trial in range(600):
        blob = bytearray(blobs[trial % len(blobs)])
        for _ in range(int(rng.integers(1, 4))):
            blob[int(rng.integers(4, len(blob)))] = int(rng.integers(0, 256)) if rng.random() < 0.5 else int(rng.integers(0, 8))
        if rng.random() < 0.1:
            blob = blob[: int(rng.integers(8, len(blob)))]
        buf = ctypes.create_string_buffer(bytes(blob), len(blob))
        h = ctypes.c_void_p()
        lib_ok = L.ge_table_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(h)) == 0
        if lib_ok:
            L.ge_table_destroy(h)
        try:
            Oracle(bytes(blob))
            ora_ok = True
        except ValueError:
            ora_ok = False
        assert lib_ok == ora_ok, "trial %d: library %s, oracle %s\n%s" % (trial, lib_ok, ora_ok, bytes(blob).hex())
        accepted += lib_ok
        rejected += not lib_ok
    assert accepted > 50 and rejected > 50


@pytest.mark.gpu
def test_accepted_mutated_blobs_run_identically(games):
    """Whatever the validators let through must run the same on the GPU (interpreter kernel) and on Oracle B."""
    import ctypes
    from game_engine_b200 import capi
    from game_engine_b200.batch import SessionBatch, Table
    from oracle.oracle import Oracle
    L = capi.lib()
    rng = np.random.default_rng(123)
    blobs = [games("werewolf-(mafia)", 8).blob, games("two-truths-and-a-lie", 4).blob, random_table(5, T.FAMILY_WEREWOLF).pack(),
             random_table(6, T.FAMILY_TTL).pack()]
    ran = 0
    for trial in range(400):
        blob = bytearray(blobs[trial % len(blobs)])
        for _ in range(int(rng.integers(1, 3))):
            blob[int(rng.integers(32, len(blob)))] = int(rng.integers(0, 6))
        blob = bytes(blob)
        buf = ctypes.create_string_buffer(blob, len(blob))
        h = ctypes.c_void_p()
        if L.ge_table_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(h)) != 0:
            continue
        L.ge_table_destroy(h)
        if blob in blobs or ran >= 40:
            continue

        class _B:
            pass
        cg = _B()
        cg.blob, cg.audience_preds = blob, {}
        cg.record_size = T.record_size(blob[6], blob[8])
        o = Oracle(blob)
        n, first, seed = 300, 17, trial
        b = SessionBatch(Table(cg), n, first_session_id=first, seed=seed, kernel="tps")
        rec = o.init(n)
        for k in range(24):
            b.step(1)
            o.step(rec, first, seed, 1)
            assert np.array_equal(b.export_state(), rec), "trial %d step %d blob %s" % (trial, k, blob.hex())
        b.close()
        ran += 1
    assert ran >= 20
