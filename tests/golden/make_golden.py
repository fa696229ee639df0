"""Generates the golden fixtures in this directory from the UNMODIFIED reference core
(oracle/_ref/libsfref.so, compiled in place from /root/reference by oracle/Makefile).
Run in the build container:  python tests/golden/make_golden.py
Fixtures (all small, committed):
  known_answers.json      rand() stream, first spawns, hexagon vertices, wireframes, episode length, dump strings
  trace_<gametype>.npz    3000 random-action steps per game type (seed 1): key masks and the reference's
                          reward / done / fort_kill / events per step + full state records every 100 steps
  trace_autoturn_kill.npz scripted double-shot policy (covers vulnerability >= 11 and fortress kills)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.oracle import OracleEnv, RefEnv, Record, record_dtype, wireframe  # noqa: E402
from conftest import scripted_kill_policy  # noqa: E402


def rec_bytes(r):
    return np.frombuffer(bytes(r), dtype=np.uint8).copy()


def trace(gametype, T, policy, seed=1):
    ref = RefEnv(gametype, seed)
    helper = OracleEnv(gametype, seed)  # only for the action->keymask table
    km = np.zeros(T, np.uint8); rew = np.zeros(T, np.int32); done = np.zeros(T, np.uint8)
    kill = np.zeros(T, np.uint8); ev = np.zeros(T, np.uint32)
    recs, rec_t = [rec_bytes(ref.get_state())], [0]
    rng = np.random.RandomState(1234)
    for t in range(T):
        if policy == "random":
            k = helper.keymask(rng.randint(helper.num_actions(1)))
        else:
            k = scripted_kill_policy(t, ref.get_state().vulnerability)
        r, d, f, e = ref.step(k)
        km[t], rew[t], done[t], kill[t], ev[t] = k, r, d, f, e
        if d:
            ref.reset()
        if (t + 1) % 100 == 0:
            recs.append(rec_bytes(ref.get_state())); rec_t.append(t + 1)
    return dict(keymask=km, reward=rew, done=done, fort_kill=kill, events=ev, records=np.stack(recs), record_t=np.array(rec_t))


def main():
    ka = {}
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    ka["rand_seed1"] = [libc.rand() for _ in range(12)]
    spawns = []
    r = RefEnv("youturn", 1)
    for _ in range(5):
        s = r.get_state(); spawns.append([s.ship_x, s.ship_y, s.ship_angle, s.rng_count]); r.reset()
    ka["spawns_seed1"] = spawns
    s = RefEnv("youturn", 12345).get_state(); ka["spawn_seed12345"] = [s.ship_x, s.ship_y, s.ship_angle]
    ka["initial_dump_seed1"] = RefEnv("youturn", 1).dump()
    big, small = RefEnv("youturn", 1).hexagons()
    ka["hex_big"] = big.tolist(); ka["hex_small"] = small.tolist()
    ka["wireframes"] = {n: wireframe(i).tolist() for i, n in enumerate(["missile", "shell", "ship", "fortress"])}
    r = RefEnv("autoturn", 1); n = 0
    while True:
        n += 1
        if r.step(0)[1]:
            break
    ka["episode_ticks"] = n; ka["episode_time_ms"] = r.get_state().time
    ka["record_sizeof"] = C.sizeof(Record)
    ka["start_vel"] = [RefEnv("youturn", 1).get_state().ship_vx, RefEnv("youturn", 1).get_state().ship_vy]
    json.dump(ka, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    for gt in ("youturn", "autoturn", "test-youturn", "test-autoturn"):
        np.savez_compressed(os.path.join(HERE, "trace_%s.npz" % gt), **trace(gt, 3000, "random"))
    np.savez_compressed(os.path.join(HERE, "trace_autoturn_kill.npz"), **trace("autoturn", 3000, "kill"))
    t = trace("autoturn", 3000, "kill")
    print("kill trace: kills", int(t["fort_kill"].sum()), "reward sum", int(t["reward"].sum()))


if __name__ == "__main__":
    main()
