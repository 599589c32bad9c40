"""The barrier protocol of the experimental three-slot forward kernel (conv_fwd_bf16.cuh, RING=1, off by default) was
written without a GPU at hand; its discrete-event model (tools/sim_fwd_ring.py) must stay deadlock- and hazard-free
under random scheduling, and must notice a broken protocol."""
import importlib.util
import os
import random


def _load():
    path = os.path.join(os.path.dirname(__file__), "..", "tools", "sim_fwd_ring.py")
    spec = importlib.util.spec_from_file_location("sim_fwd_ring", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ring_protocol_model_is_clean():
    assert _load().main(trials=600, seed=3) == 0


def test_model_notices_a_missing_handshake():
    sim = _load()

    class Broken(sim.Sim):
        def back(self):                     # a back end that never releases its accumulator slots
            for ev in super().back():
                yield ev
                if ev[0] == "step":
                    return

    assert Broken([208, 208, 208], 4, random.Random(1)).run(max_steps=20000) is not None
