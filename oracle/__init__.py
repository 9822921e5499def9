"""CPU oracles for the referee/phase step — TEST INFRASTRUCTURE ONLY (see ge_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  game_engine_b200 never does.
"""
