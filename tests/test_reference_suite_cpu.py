"""CPU dry run of the restated reference suite (tests/test_gpu_reference_suite.py): the fp64 oracle stands in for the CUDA
engine behind `LEFTNetB200.forward`, everything else — the reference's fixtures, this package's `EGNNDynamics` host path
(ragged / empty fragments, 1-element integer `t`, ignored `edge_attr`, shared encoders assigned after construction), the
properties and their thresholds — is exactly what runs on the GPU.  It validates the TESTS and the host modules; the
kernels are judged by the GPU run."""
import pytest
import torch

import oareactdiff_b200 as ob
import tests.test_gpu_reference_suite as suite
from oracle import oa_ref


def _oracle_forward(self, h, pos, edge_index, edge_attr=None, node_mask=None, edge_mask=None, update_coords_mask=None,
                    subgraph_mask=None):
    sd = {k: v.detach().double() for k, v in self.state_dict().items()}
    ho, dpos = oa_ref.leftnet_forward(sd, dict(self.cfg), h.double(), pos.double(), edge_index,
                                      subgraph_mask if self.cfg["object_aware"] else None)
    return ho.to(h.dtype), pos + dpos.to(pos.dtype), None


CASES = [(n, ()) for n in sorted(dir(suite)) if n.startswith("test_") and n != "test_switch_fragments"]
CASES += [("test_switch_fragments", (False,)), ("test_switch_fragments", (True,))]


@pytest.mark.parametrize("name,args", CASES, ids=[f"{n}{list(a) if a else ''}" for n, a in CASES])
def test_restated_reference_suite_holds_for_the_oracle(monkeypatch, name, args):
    monkeypatch.setattr(suite, "DEV", torch.device("cpu"))
    monkeypatch.setattr(ob.LEFTNetB200, "forward", _oracle_forward)
    monkeypatch.setattr(ob.EGNNDynamics, "fused_ok", lambda self, device: False)
    with torch.no_grad():
        getattr(suite, name)(*args)
