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
