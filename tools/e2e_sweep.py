"""End-to-end (host buffers, public C ABI) throughput of the flood for several partition counts / host threads."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--groups", type=int, default=100_000)
ap.add_argument("--members", type=int, default=5)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--engines", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--threads", type=int, nargs="+", default=[16])
ap.add_argument("--drivers", type=int, nargs="+", default=[0], help="driver threads (0 = the default rule)")
a = ap.parse_args()
for T, D in [(t, d) for t in a.threads for d in a.drivers]:
    os.environ["RA_HOSTSIM_THREADS"] = str(T)
    if D: os.environ["RA_HOSTSIM_DRIVERS"] = str(D)
    else: os.environ.pop("RA_HOSTSIM_DRIVERS", None)
    from ra_b200.engine import Engine, HostFlood
    for K in a.engines:
        G, M = a.groups, a.members
        engs = [Engine(G // K + (1 if i < G % K else 0), M, route_on_device=True) for i in range(K)]
        for e in engs: e.reset_empty()
        hf = HostFlood(engs)
        hf.run(0, 1, 10, seed=0xA00, bootstrap=True)
        hf.run(40, 1, 10, seed=0xA00)
        c0 = sum(e.counters()["commits"] for e in engs)
        st = hf.run(a.steps, 1, 10, seed=0xA00)
        c1 = sum(e.counters()["commits"] for e in engs)
        print("engines %d threads %d drivers %s: %.1f M commits/s  %.3f ms/step  (wait %.3f model %.3f ms/step)  h2d %.1f MB d2h %.1f MB per step"
              % (K, T, D or "auto", (c1 - c0) / st["seconds"] / 1e6, st["seconds"] * 1e3 / a.steps, st["step_seconds"] * 1e3 / a.steps,
                 st["model_seconds"] * 1e3 / a.steps, st["h2d_bytes"] / a.steps / 1e6, st["d2h_bytes"] / a.steps / 1e6), flush=True)
        hf.close()
        for e in engs: e.close()
