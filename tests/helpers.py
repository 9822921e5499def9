"""Shared test helpers: replaying packed records through the host adapter and normalising dict states."""
import copy


KEEP = ("current_phase_id", "current_phase_name", "player_states", "playerActions", "phase_history", "game_notes")


def normalise(state):
    s = copy.deepcopy({k: state.get(k) for k in KEEP})
    for e in s.get("phase_history") or []:
        e.pop("timestamp", None)
    for pa in (s.get("playerActions") or {}).values():
        for a in (pa.get("actions") or {}).values():
            a.pop("timestamp", None)
    s["current_phase_name"] = s.get("current_phase_name") or ""
    return s


def replay_records(cg, records):
    """records[k] = canonical record after k steps (records[0] = initial).  Returns the dict trace the host
    adapter materialises from them: trace[k] comparable with Oracle A's trace[k]."""
    from game_engine_b200.adapter import SessionCodec
    codec = SessionCodec(cg)
    state = codec.initial_state()
    trace = [normalise(state)]
    for k in range(1, len(records)):
        upd = codec.step_update(state, records[k - 1], records[k], now_ms=0, now_iso="")
        state.update(upd)
        trace.append(normalise(state))
    return trace


def oracle_b_records(o, sid, seed, n_steps):
    rec = o.init(1)
    out = [rec[0].copy()]
    for _ in range(n_steps):
        o.step(rec, sid, seed, 1)
        out.append(rec[0].copy())
    return out


def first_diff(a, b, path=""):
    """Human-readable location of the first difference between two JSON-like values."""
    if type(a) != type(b):
        return "%s: %r != %r" % (path, a, b)
    if isinstance(a, dict):
        for k in sorted(set(a) | set(b), key=str):
            if k not in a or k not in b:
                return "%s/%s: missing on one side (%r vs %r)" % (path, k, a.get(k), b.get(k))
            d = first_diff(a[k], b[k], path + "/" + str(k))
            if d:
                return d
        return None
    if isinstance(a, list):
        if len(a) != len(b):
            return "%s: len %d != %d (%r vs %r)" % (path, len(a), len(b), a[-1:], b[-1:])
        for i, (x, y) in enumerate(zip(a, b)):
            d = first_diff(x, y, "%s[%d]" % (path, i))
            if d:
                return d
        return None
    return None if a == b else "%s: %r != %r" % (path, a, b)


def pick(paths, *needles):
    """Fixtures whose file name contains one of the needles, in the order given (stable under new fixtures)."""
    out = []
    for nd in needles:
        hit = [p for p in paths if nd in p.rsplit("/", 1)[-1]]
        assert len(hit) == 1, (nd, hit)
        out.append(hit[0])
    return out


def play_with_human(cg, messages, step_fn, human_seats=(1,)):
    """Replays a game with a person in the given seats.  messages[k] = what they typed before step k + 1.
    step_fn(record_before, human_mask, inputs_row) -> record_after does the stepping (Oracle B or the GPU).
    Per step: the router's logging of the message (adapter.log_human_action), the inputs the adapter reads from it,
    the step, and the adapter's update.  Returns (records, dict trace) like replay_records."""
    from game_engine_b200.adapter import SessionCodec
    codec = SessionCodec(cg)
    state = codec.initial_state()
    rec = codec.record_from_state(state)
    records, trace = [rec.copy()], [normalise(state)]
    for text in messages:
        state["messages"] = [{"type": "human", "content": text}]
        state["playerActions"] = codec.log_human_action(state, text, now_ms=0)
        mask, row = codec.human_inputs(state, human_seats)
        before = codec.record_from_state(state)
        after = step_fn(before, mask, row)
        state.update(codec.step_update(state, before, after, now_ms=0, now_iso="", human_mask=mask))
        records.append(after.copy())
        trace.append(normalise(state))
    return records, trace


def oracle_step_fn(o, sid, seed):
    import numpy as np

    def step(before, mask, row):
        rec = np.array(before.reshape(1, -1), dtype=np.uint8, copy=True)
        o.step_humans(rec, sid, seed, np.array([mask], dtype=np.uint32), row.reshape(1, -1))
        return rec[0]
    return step
