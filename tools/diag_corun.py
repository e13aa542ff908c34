"""Diagnostic: do kernels launched from two Python threads on two streams run side by side on this box, and does the
virtual-rank peer exchange work?  Prints timings and the library's own error text."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

torch.cuda.set_device(0)
torch.cuda.init()
x = torch.zeros(1, device="cuda")
cyc = int(1.0e9)  # ~0.5 s at 1.9 GHz


def spin(res, i):
    with torch.cuda.stream(torch.cuda.Stream()):
        torch.cuda._sleep(cyc)
        torch.cuda.current_stream().synchronize()
    res[i] = time.perf_counter()


t0 = time.perf_counter()
res = [0, 0]
ts = [threading.Thread(target=spin, args=(res, i)) for i in range(2)]
[t.start() for t in ts]
[t.join() for t in ts]
print("two spin kernels from two threads: %.2f s (one alone ~%.2f s)" % (max(res) - t0, cyc / 1.9e9), flush=True)

os.environ["SSDR_PEER_TIMEOUT_MS"] = "3000"
from ssdr_al_b200 import device as D, dist as SD
F = torch.randn((60_000, 32), device="cuda")
want = D.fps(F, 50, 11)
sms = torch.cuda.get_device_properties(0).multi_processor_count
T0 = time.perf_counter()


def attempt(tag, world, warm, delay):
    groups = SD.PeerGroup.local(world)
    gate = threading.Barrier(world)

    def work(r):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                if warm:  # creates this thread's context, streams and workspaces before the ranks depend on each other
                    D.fps(F, 5, 11)
                    torch.cuda.current_stream().synchronize()
                gate.wait()
                time.sleep(delay * r)
                t1 = time.perf_counter()
                print("%s rank %d enters at %.3f" % (tag, r, t1 - T0), flush=True)
                o = SD.fps_sharded(F, 50, 11, groups[r], max_ctas=sms // world)
                torch.cuda.current_stream().synchronize()
                print("%s rank %d ok after %.3f s equal=%s" % (tag, r, time.perf_counter() - t1, bool(torch.equal(o, want))), flush=True)
        except Exception as e:  # noqa: BLE001
            print("%s rank %d FAILED after %.3f s: %s" % (tag, r, time.perf_counter() - t1, e), flush=True)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for g in groups:
        g.destroy()


def attempt_deferred(tag, world):
    """All ranks enqueue first (SSDR_PEER_DEFER_CHECK=1), meet at a host barrier, then wait for their streams."""
    os.environ["SSDR_PEER_DEFER_CHECK"] = "1"
    groups = SD.PeerGroup.local(world)
    gate = threading.Barrier(world)

    def work(r):
        t1 = time.perf_counter()
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                D.fps(F, 5, 11)
                torch.cuda.current_stream().synchronize()
                gate.wait()
                t1 = time.perf_counter()
                o = SD.fps_sharded(F, 50, 11, groups[r], max_ctas=sms // world)
                gate.wait()
                groups[r].check()
                print("%s rank %d ok after %.3f s equal=%s" % (tag, r, time.perf_counter() - t1, bool(torch.equal(o, want))), flush=True)
        except Exception as e:  # noqa: BLE001
            print("%s rank %d FAILED after %.3f s: %s" % (tag, r, time.perf_counter() - t1, e), flush=True)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for g in groups:
        g.destroy()
    os.environ.pop("SSDR_PEER_DEFER_CHECK", None)


attempt_deferred("deferred2", 2)
attempt_deferred("deferred4", 4)
attempt("cold", 2, False, 0.0)
