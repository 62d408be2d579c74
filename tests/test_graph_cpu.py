"""GraphRunner's control flow (eager -> capture -> replay; a refused capture degrades to eager) with torch.cuda mocked out."""
import torch

from parakeet_b200 import graph as G


def test_graph_runner_state_machine(monkeypatch):
    class FakeGraph:
        replays = 0

        def replay(self):
            FakeGraph.replays += 1

    class Ctx:
        def __init__(self, fail):
            self.fail = fail

        def __enter__(self):
            if self.fail:
                raise RuntimeError("capture refused")

        def __exit__(self, *a):
            return False

    fail = {"v": False}
    monkeypatch.setattr(torch.cuda, "synchronize", lambda: None)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", lambda g: Ctx(fail["v"]))
    r = G.GraphRunner()
    r.enabled = True
    calls = []
    fn = lambda x: (calls.append(1), x * 2)[1]
    x = torch.ones(3)
    a = r.run("k", fn, [x])                       # first sight of the key: eager
    r.run("k", fn, [x])                           # second: capture (fn runs once more under the fake capture) + replay
    r.run("k", fn, [x + 1])                       # third: replay only
    assert len(calls) == 2 and FakeGraph.replays == 2 and torch.equal(a, x * 2) and r.replays == 2
    fail["v"] = True
    for _ in range(3):
        out = r.run("j", fn, [x])
    assert "j" in r._disabled and torch.equal(out, x * 2) and FakeGraph.replays == 2   # refused capture: eager from then on
    r.clear()
    assert not r._graphs and not r._disabled and not r._seen
    r.enabled = False
    assert torch.equal(r.run("k", fn, [x]), x * 2) and not r._seen                     # PK_CUDA_GRAPHS=0: always eager


def test_zero_planes_lru_is_tied_to_the_graphs():
    """training/wgrad.py ZeroPlanes: persistent zero-initialised operand planes are filed per batch geometry, bounded (LRU), and
    evicting a geometry tells the owner to drop the CUDA graph captured for it (its buffer addresses are baked into the graph)."""
    import torch
    from parakeet_b200.graph import GraphRunner
    from parakeet_b200.training.wgrad import ZeroPlanes
    runner = GraphRunner(max_graphs=8)
    dropped = []

    def on_evict(key):
        dropped.append(key)
        runner.drop(key)

    zp = ZeroPlanes(max_geoms=2, on_evict=on_evict)
    zp.begin("a")
    pa = zp.get(("xt", 2, 10), (4, 64), "cpu")
    assert pa.hi.shape == (4, 64) and not pa.hi.any() and not pa.lo.any()
    assert zp.get(("xt", 2, 10), (4, 64), "cpu") is pa                    # same key, same geometry: the same buffer
    assert zp.get(("dyt", 2, 10), (4, 64), "cpu") is not pa               # the role keeps live operands apart
    pa.hi.fill_(1)                                                         # "valid region" written by a transpose
    zp.begin("b")
    assert zp.get(("xt", 2, 10), (4, 64), "cpu") is not pa                # other geometry: its own planes (its own padding)
    zp.touch("a")                                                          # a graph replay of "a" keeps it recent
    runner._seen.add("b")
    zp.begin("c")                                                          # bound 2: "b" (least recently used) goes, with its graph
    assert dropped == ["b"] and "b" not in runner._seen and len(zp) == 2
    zp.begin("a")
    assert zp.get(("xt", 2, 10), (4, 64), "cpu") is pa                    # "a" survived
    zp.begin("b")                                                          # back again: fresh zeros
    assert dropped == ["b", "c"] and not zp.get(("xt", 2, 10), (4, 64), "cpu").hi.any()
