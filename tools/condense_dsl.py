#!/usr/bin/env python3
"""Condense a reference game DSL file to the keys the transition-table compiler consumes.

The GPU box has no /root/reference, so the game definitions (input DATA of the path, in the
reference's own grammar: prompt/dsl_phases_generation_prompt.txt:84-149) ship with the package as
game_engine_b200/games/*.yaml.  UI action lists, prose descriptions and example states are
dropped; phase ids, names, completion criteria, target-player conditions, next_phase forms, role
names, the state template and audience-group criteria are kept verbatim.

tests/test_compiler.py checks (when /root/reference exists) that compiling the reference file and
the condensed file gives byte-identical tables.

usage: python tools/condense_dsl.py /root/reference/games game_engine_b200/games
       python tools/condense_dsl.py "/root/reference/game_draft/werewolf-(mafia).yaml" game_engine_b200/games/werewolf-draft.yaml
(the second form condenses ONE file under another name: the reference's earlier 13-phase werewolf generation,
 which no reference code loads — utils.py:565 reads games/ only — and which serves here as a third table with a
 different phase graph, two terminal phases and a different per-player state schema)
"""
import sys
import os
import yaml


def condense(dsl: dict) -> dict:
    decl = dsl.get("declaration", {})
    out_decl = {
        "is_multiplayer": decl.get("is_multiplayer", True),
        "min_players": decl.get("min_players"),
    }
    if decl.get("roles"):
        out_decl["roles"] = [{"name": r["name"]} for r in decl["roles"]]
    out_decl["player_states"] = {k: {"type": v.get("type")} for k, v in decl.get("player_states", {}).items()}
    out_decl["player_states_template"] = decl.get("player_states_template")
    if decl.get("audience_groups"):
        out_decl["audience_groups"] = {
            k: {"selection_criteria": v.get("selection_criteria")} for k, v in decl["audience_groups"].items()
        }
    phases = {}
    for pid, ph in dsl.get("phases", {}).items():
        cc = ph.get("completion_criteria", {})
        occ = {"type": cc.get("type")}
        if "wait_for" in cc:
            occ["wait_for"] = cc["wait_for"]
        if "target_players" in cc:
            occ["target_players"] = {"condition": cc["target_players"].get("condition")}
        phases[pid] = {"name": ph.get("name"), "completion_criteria": occ, "next_phase": ph.get("next_phase")}
    return {"declaration": out_decl, "phases": phases}


def main() -> None:
    src, dst = sys.argv[1], sys.argv[2]
    if os.path.isfile(src):
        with open(src, encoding="utf-8") as f:
            dsl = yaml.safe_load(f)
        with open(dst, "w", encoding="utf-8") as f:
            f.write("# condensed by tools/condense_dsl.py from the reference's %s\n" % os.path.relpath(src, "/root/reference"))
            yaml.safe_dump(condense(dsl), f, sort_keys=False, allow_unicode=True, width=100)
        print("wrote", dst)
        return
    os.makedirs(dst, exist_ok=True)
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".yaml"):
            continue
        with open(os.path.join(src, fn), encoding="utf-8") as f:
            dsl = yaml.safe_load(f)
        with open(os.path.join(dst, fn), "w", encoding="utf-8") as f:
            f.write("# condensed by tools/condense_dsl.py from the reference's games/%s\n" % fn)
            yaml.safe_dump(condense(dsl), f, sort_keys=False, allow_unicode=True, width=100)
        print("wrote", os.path.join(dst, fn))


if __name__ == "__main__":
    main()
