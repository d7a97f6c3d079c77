"""TEST INFRASTRUCTURE ONLY -- fp32 PyTorch restatement of the fork's per-sample-graph network ``gwnet_diff_G``
(model.py:244-407) and its operators ``nconv2`` / ``gcn2`` (model.py:16-22, 57-80).  Functional and state-dict driven like
oracle/gwnet_oracle.py; pinned to tests/golden/diffg.npz, which tests/tools/make_golden_diffg.py generates from the real
reference.  Imported by tests/ only."""
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def nconv2(x, A):                                            # model.py:20-22
    return torch.einsum("ncvl,nvw->ncwl", x, A).contiguous()


def gcn2(x, supports, weight, bias, order, dropout, training):   # model.py:66-80
    out = [x]
    for a in supports:
        x1 = nconv2(x, a)
        out.append(x1)
        for _ in range(2, order + 1):
            x2 = nconv2(x1, a)
            out.append(x2)
            x1 = x2
    h = F.conv2d(torch.cat(out, dim=1), weight, bias)
    return F.dropout(h, dropout, training=training)


def draw_node_embeddings(batch, num_nodes, rank=10):
    """model.py:324-329: fresh embeddings every forward, drawn on the CPU generator (nodevec1 first)."""
    return torch.randn(batch, num_nodes, rank), torch.randn(batch, rank, num_nodes)


def dilations(blocks, layers, base=4):                       # model.py:271-293
    out = []
    for _ in range(blocks):
        d = base
        for _ in range(layers):
            out.append(d)
            d *= 2
    return out


def receptive_field(blocks, layers, kernel_size=2):          # model.py:270-294 (computed as if dilations were 1, 2, ...)
    rf = 1
    for _ in range(blocks):
        scope = kernel_size - 1
        for _ in range(layers):
            rf += scope
            scope *= 2
    return rf


def forward(state: Dict[str, torch.Tensor], inp, supports: Optional[Sequence[torch.Tensor]], nodevecs, *, blocks=4, layers=2,
            gcn_bool=True, addaptadj=True, dropout=0.0, training=True, order=2, momentum=0.1, eps=1e-5):
    """``gwnet_diff_G.forward(input, supports, aptinit=None)`` (model.py:313-407) with the node embeddings given."""
    rf = receptive_field(blocks, layers)
    T = inp.size(3)
    x = F.pad(inp, (rf - T, 0, 0, 0)) if T < rf else inp      # :337-341
    x = F.conv2d(x, state["start_conv.weight"], state["start_conv.bias"])
    if gcn_bool and addaptadj and supports is None:
        supports = []
    new_supports = None
    if gcn_bool and addaptadj and supports is not None:       # :345-347
        adp = F.softmax(F.relu(torch.matmul(nodevecs[0], nodevecs[1])), dim=2)
        new_supports = list(supports) + [adp]
    skip = None
    for i, d in enumerate(dilations(blocks, layers)):
        residual = x
        f = torch.tanh(F.conv2d(residual, state[f"filter_convs.{i}.weight"], state[f"filter_convs.{i}.bias"], dilation=(1, d)))
        g = torch.sigmoid(F.conv2d(residual, state[f"gate_convs.{i}.weight"], state[f"gate_convs.{i}.bias"], dilation=(1, d)))
        x = f * g
        s = F.conv2d(x, state[f"skip_convs.{i}.weight"], state[f"skip_convs.{i}.bias"])
        skip = s if skip is None else s + skip[:, :, :, -s.size(3):]
        if gcn_bool and supports is not None:                 # :388-394
            sup = new_supports if addaptadj else supports
            x = gcn2(x, sup, state[f"gconv.{i}.mlp.mlp.weight"], state[f"gconv.{i}.mlp.mlp.bias"], order, dropout, training)
        else:
            x = F.conv2d(x, state[f"residual_convs.{i}.weight"], state[f"residual_convs.{i}.bias"])
        x = x + residual[:, :, :, -x.size(3):]
        x = F.batch_norm(x, state[f"bn.{i}.running_mean"], state[f"bn.{i}.running_var"], state[f"bn.{i}.weight"],
                         state[f"bn.{i}.bias"], training, momentum, eps)
        if training:
            state[f"bn.{i}.num_batches_tracked"] += 1
    x = F.relu(skip)
    x = F.relu(F.conv2d(x, state["end_conv_1.weight"], state["end_conv_1.bias"]))
    return F.conv2d(x, state["end_conv_2.weight"], state["end_conv_2.bias"])
