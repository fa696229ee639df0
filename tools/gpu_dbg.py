"""Small debug driver: a few envs, a few steps, frames compared with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spacefortress_b200 import SFVecEnv
from oracle.oracle import OracleEnv
gt = sys.argv[1] if len(sys.argv) > 1 else "youturn"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
seeds = np.arange(1, n + 1)
env = SFVecEnv(gt, num_envs=n, device=0, seeds=seeds)
obs = env.reset()
orc = [OracleEnv(gt, int(s)) for s in seeds]
bad = 0
for i in range(n):
    if not np.array_equal(orc[i].obs(), obs[i, 0]):
        d = np.argwhere(orc[i].obs() != obs[i, 0]); print("reset frame differs env", i, len(d), d[:5]); bad += 1
rng = np.random.RandomState(3)
for t in range(T):
    a = rng.randint(0, env.num_actions, size=n)
    try:
        obs, rew, done, info = env.step(a)
    except Exception as ex:
        print('FAIL at t', t, ex)
        for i in range(n):
            r, d, k, e = orc[i].step(orc[i].keymask(int(a[i])))
            print(i, 'ev %x' % e, orc[i].dump())
        sys.exit(1)
    for i in range(n):
        r, d, k, e = orc[i].step(orc[i].keymask(int(a[i])))
        o = orc[i].obs()
        if not np.array_equal(o, obs[i, 0]):
            dd = np.argwhere(o != obs[i, 0])
            if bad < 12:
                print("t", t, "env", i, "npx", len(dd), "first", dd[:4].tolist(), "exp", o[tuple(dd[0])], "got", obs[i, 0][tuple(dd[0])], "ev %x" % e)
            bad += 1
print("mismatching frames:", bad, "of", n * T)
