"""The one parity-unpinned operator of the path: ``torch_geometric.nn.GraphSAGE`` (reference userEncoders.py:6,54-58,153).
PyG is not installed in the build image and the reference pins no version, so oracle/ref_import.py and the CUDA kernels
restate the published SAGEConv(mean) algorithm.  Wherever a real torch_geometric IS importable this test closes the gap:
it runs PyG's own GraphSAGE(400, 400, num_layers=1) on the reference's bipartite edge_index and compares it with the
restatement (same weights).  Skipped otherwise (mitigation ii)."""
import importlib.util

import pytest
import torch

from oracle import lime_oracle as O
from oracle import ref_import

pyg_missing = importlib.util.find_spec("torch_geometric") is None


def _edge_index(num_users, max_history_num):
    """userEncoders.CROWN.create_bipartite_graph (userEncoders.py:91-98): source = user node u (index < num_users),
    target = every history slot."""
    row = torch.arange(num_users).view(-1, 1).repeat(1, max_history_num).view(-1)
    col = torch.arange(max_history_num).view(1, -1).repeat(num_users, 1).view(-1)
    return torch.stack([row, col], dim=0)


@pytest.mark.skipif(pyg_missing, reason="torch_geometric is not installed (parity of GraphSAGE stays unpinned here)")
@pytest.mark.parametrize("B,H,bs", [(8, 50, 8), (64, 50, 64), (5, 50, 32)])
def test_graphsage_restatement_matches_real_pyg(B, H, bs):
    from torch_geometric.nn import GraphSAGE
    torch.manual_seed(0)
    D = 400
    real = GraphSAGE(in_channels=D, hidden_channels=D, num_layers=1, out_channels=D, dropout=0.0).double().eval()
    x = torch.randn(B, H + bs, D, dtype=torch.float64)
    with torch.no_grad():
        want = real(x, _edge_index(B, H))
    sd = {O.USER + "graph_sage.convs.0.lin_l.weight": real.convs[0].lin_l.weight.detach(),
          O.USER + "graph_sage.convs.0.lin_l.bias": real.convs[0].lin_l.bias.detach(),
          O.USER + "graph_sage.convs.0.lin_r.weight": real.convs[0].lin_r.weight.detach()}
    got = O.graph_sage(sd, x, B, H, torch.float64)
    assert torch.allclose(got[:, :H], want[:, :H], rtol=1e-10, atol=1e-10)
    # and the import stub the oracle pins use
    stub = ref_import._GraphSAGE(D, D, 1, out_channels=D).double()
    stub.load_state_dict({k.replace(O.USER + "graph_sage.", ""): v for k, v in sd.items()})
    with torch.no_grad():
        assert torch.allclose(stub(x, _edge_index(B, H))[:, :H], want[:, :H], rtol=1e-10, atol=1e-10)


def test_stub_and_oracle_agree_without_pyg():
    """Always runs: the two restatements (import stub used to run the reference, closed-form oracle) agree, including the
    runtime batch > H case where user-node rows enter the mean."""
    torch.manual_seed(1)
    D, H = 400, 50
    for B, bs in ((8, 8), (64, 64)):
        stub = ref_import._GraphSAGE(D, D, 1, out_channels=D).double()
        x = torch.randn(B, H + bs, D, dtype=torch.float64)
        sd = {O.USER + "graph_sage." + k: v.detach() for k, v in stub.state_dict().items()}
        with torch.no_grad():
            want = stub(x, _edge_index(B, H))
        got = O.graph_sage(sd, x, B, H, torch.float64)
        assert torch.allclose(got[:, :H], want[:, :H], rtol=1e-10, atol=1e-10)
