"""First-contact GPU check with verbose diagnostics (writes gpurun_out/first_check.txt)."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "first_check.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a)
    print(s); log.write(s + "\n"); log.flush()
try:
    import torch
    from spacefortress_b200 import SFVecEnv
    from oracle.oracle import OracleEnv, Record, draw_native, draw_obs
    import ctypes as C
    P("torch", torch.__version__, torch.cuda.get_device_name(0))
    for gametype in ("youturn", "autoturn"):
        n, T = 32, 600
        seeds = np.arange(1, n + 1)
        env = SFVecEnv(gametype, num_envs=n, device=0, seeds=seeds)
        t0 = time.time(); obs = env.reset(); P(gametype, "reset ok", time.time() - t0, obs.shape)
        orc = [OracleEnv(gametype, int(s)) for s in seeds]
        nat = env.render_frames(native=True)
        bad = 0
        for i in range(n):
            o = orc[i].native_frame()
            d = np.abs(o.astype(int) - nat[i])
            if d.max() > 0:
                bad += 1
                if bad <= 3:
                    ys, xs = np.nonzero(d); P(" init native diff env", i, "npx", len(ys), "max", d.max(), "bbox", ys.min(), ys.max(), xs.min(), xs.max())
            if not np.array_equal(orc[i].obs(), obs[i, 0]): P(" init obs84 differs env", i, np.abs(orc[i].obs().astype(int) - obs[i, 0]).max())
        P(gametype, "initial frames: envs with native diffs:", bad)
        rng = np.random.RandomState(0)
        nstep_bad = nstate_bad = nframe_bad = 0
        frame_checks = 0
        for t in range(T):
            a = rng.randint(0, env.num_actions, size=n)
            obs, rew, done, info = env.step(a)
            ev = env.last_events
            for i in range(n):
                r, d, k, e = orc[i].step(orc[i].keymask(int(a[i])))
                if (r, d, k, e) != (int(rew[i]), bool(done[i]), bool(info[i]), int(ev[i])):
                    nstep_bad += 1
                    if nstep_bad <= 5: P(" step mismatch t", t, "env", i, "oracle", (r, d, k, hex(e)), "gpu", (int(rew[i]), bool(done[i]), bool(info[i]), hex(int(ev[i]))))
            if t % 25 == 0 or t == T - 1:
                recs = env.get_state()
                nat = env.render_frames(native=True)
                for i in range(n):
                    so = orc[i].get_state()
                    g = recs[i]
                    for f in Record.INT_FIELDS:
                        if int(getattr(so, f)) != int(getattr(g, f)):
                            nstate_bad += 1
                            if nstate_bad <= 10: P(" state mismatch t", t, "env", i, f, int(getattr(so, f)), int(getattr(g, f)))
                    if list(so.stats) != list(g.stats):
                        nstate_bad += 1; P(" stats mismatch", t, i, list(so.stats), list(g.stats))
                    for f in ("ship_x", "ship_y", "ship_vx", "ship_vy", "ship_angle", "fortress_angle", "fortress_last_angle", "points", "raw_points"):
                        if float(getattr(so, f)) != float(getattr(g, f)):
                            nstate_bad += 1
                            if nstate_bad <= 10: P(" float mismatch t", t, "env", i, f, repr(float(getattr(so, f))), repr(float(getattr(g, f))))
                    frame_checks += 1
                    o = orc[i].native_frame()
                    dd = np.abs(o.astype(int) - nat[i])
                    if dd.max() > 0:
                        nframe_bad += 1
                        if nframe_bad <= 8:
                            ys, xs = np.nonzero(dd)
                            P(" native frame diff t", t, "env", i, "npx", len(ys), "max", dd.max(), "bbox y", ys.min(), ys.max(), "x", xs.min(), xs.max(),
                              "ship_alive", so.ship_alive, "mm", hex(so.missile_mask), "sm", hex(so.shell_mask), "fort", so.fortress_alive)
                    if not np.array_equal(orc[i].obs(), obs[i, 0]):
                        nframe_bad += 1
                        if nframe_bad <= 8: P(" obs84 diff t", t, "env", i, np.abs(orc[i].obs().astype(int) - obs[i, 0]).max())
        P(gametype, "steps bad", nstep_bad, "state bad", nstate_bad, "frames bad", nframe_bad, "of", frame_checks)
        env.close()
    # quick throughput probe
    for n in (4096, 65536):
        env = SFVecEnv("autoturn", num_envs=n, device=0)
        env.reset(to_numpy=False)
        out = env.rollout(8)
        torch.cuda.synchronize()
        for T in (16,):
            obs = torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda")
            outd = {"obs": obs}
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record(); env.rollout(T, out=outd); e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e)
            P("rollout n", n, "T", T, "ms", ms, "steps/s", n * T / ms * 1e3, "GB/s", n * T * 8362 / ms / 1e6)
        a = torch.zeros(n, dtype=torch.int32, device="cuda")
        for _ in range(3): env.step(a)
        torch.cuda.synchronize()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): env.step(a)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        P("step n", n, "ms/step", ms, "steps/s", n / ms * 1e3)
        env2 = SFVecEnv("autoturn", num_envs=n, device=0, render=False)
        env2.reset()
        env2.rollout(8); torch.cuda.synchronize()
        s.record(); env2.rollout(64); e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        P("state-only rollout n", n, "T 64 ms", ms, "steps/s", n * 64 / ms * 1e3)
        env.close(); env2.close()
except Exception:
    P("EXCEPTION", traceback.format_exc())
    sys.exit(1)
