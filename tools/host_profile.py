"""Where does the HOST time of a step go?  cProfile over K steps of a bench preset (no syncs inside the loop), top
functions by cumulative time.  Not a test.   python tools/host_profile.py --preset with_angle_and_depth --view 256x341"""
import argparse, cProfile, io, os, pstats, sys, tempfile, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--preset", default="only2D")
ap.add_argument("--view", default="480x640")
ap.add_argument("--texture", type=int, default=2048)
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
sys.argv = ["bench.py", "--preset", a.preset, "--view", a.view, "--texture", str(a.texture)]
args = bench.parse_args()
dev = torch.device("cuda", 0)
real = sys.stdout
sys.stdout = sys.stderr
mdl, opt = bench.build_ours(args, dev, tempfile.mkdtemp())
views = [v.to(dev).as_batch() for v in bench.make_views(args, 0)]
mdl.cache_view_plans = True
for i in range(8):
    bench.one_step(mdl, opt, views[i % len(views)], i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(a.steps):
    bench.one_step(mdl, opt, views[i % len(views)], i)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / a.steps
# host cost of ENQUEUEING a step into an empty launch queue (in a free-running loop the queue fills up and every
# launch blocks on the GPU, so the loop's host time equals the device time)
host = 0.0
for i in range(a.steps):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bench.one_step(mdl, opt, views[i % len(views)], i)
    host += time.perf_counter() - t1
host /= a.steps
pr = cProfile.Profile()
for i in range(a.steps):
    torch.cuda.synchronize()
    pr.enable()
    bench.one_step(mdl, opt, views[i % len(views)], i)
    pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
sys.stdout = real
print(f"host enqueue {host * 1e3:.3f} ms/step, wall {wall * 1e3:.3f} ms/step ({a.preset} {a.view} tex {a.texture})")
print(s.getvalue()[:6000])
