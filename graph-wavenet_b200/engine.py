"""Mirror of the reference ``engine.trainer`` for the gwnet path (engine.py:9-58,119-130): same
constructor arguments, attributes (``model optimizer scaler clip loss``) and ``train`` / ``eval``
signatures and return values.  The forward/backward inside runs on the native CUDA plan."""
import torch
import torch.nn as nn
import torch.optim as optim

if __package__:
    from . import metrics as util
    from . import fused as _fused
    from .model import gwnet, gwnet_diff_G
else:  # top-level import next to ``model`` (the reference's style)
    import importlib as _il
    from model import gwnet, gwnet_diff_G  # noqa: F401
    util = _il.import_module("graph_wavenet_b200.metrics")
    _fused = _il.import_module("graph_wavenet_b200.fused")


class trainer():
    def __init__(self, scaler, in_dim, seq_length, num_nodes, nhid, dropout, lrate, wdecay, device, supports, gcn_bool,
                 addaptadj, aptinit, blocks=4, layers=2):
        per_sample = type(supports) == dict
        if per_sample:   # engine.py:14-25 -- a different graph for every sample: {'train': [..], 'val': [..], 'test': [..]}
            supports_len = 0
            for k in supports:
                supports_len = len(supports[k])
                break
            if gcn_bool and addaptadj:
                supports_len += 1
            self.model = gwnet_diff_G(device, num_nodes, dropout, supports_len, gcn_bool=gcn_bool, addaptadj=addaptadj,
                                      in_dim=in_dim, out_dim=seq_length, residual_channels=nhid, dilation_channels=nhid,
                                      skip_channels=nhid * 8, end_channels=nhid * 16, blocks=blocks, layers=layers)
        else:
            self.model = gwnet(device, num_nodes, dropout, supports=supports, gcn_bool=gcn_bool, addaptadj=addaptadj,
                               aptinit=aptinit, in_dim=in_dim, out_dim=seq_length, residual_channels=nhid,
                               dilation_channels=nhid, skip_channels=nhid * 8, end_channels=nhid * 16, blocks=blocks,
                               layers=layers)
        self.model.to(device)
        # engine.py:33 -- Adam(lr, weight_decay as L2-in-gradient).  The fused step (default) runs clip + Adam as one
        # pass of gwn_adam_step over flat buffers; GWNET_B200_FUSED_STEP=0 keeps torch.optim.Adam on the autograd path.
        self.fused = _fused.fused_enabled() and not per_sample     # the captured step is the gwnet / trainer.train path
        self.use_graph = _fused.graph_enabled()
        if self.fused:
            self.optimizer = _fused.FusedAdam(self.model.parameters(), lr=lrate, weight_decay=wdecay)
        else:
            self.optimizer = optim.Adam(self.model.parameters(), lr=lrate, weight_decay=wdecay)
        self._steps = {}
        self.loss = util.masked_mae
        self.scaler = scaler
        self.clip = 5
        self.supports = supports
        self.aptinit = aptinit
        self.state = None
        self.world = 1
        self.rank = 0
        self.p2p = False
        self._p2p_comm = None

    # ---- data parallelism (no reference counterpart): one process per GPU, batch sharded,
    # parameters replicated, ONE NCCL all-reduce per step over the flat gradient buffer the backward call fills.
    def enable_data_parallel(self):
        import torch.distributed as dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        # gradient exchange of the fused step: a peer-memory kernel over NVLink (fused.P2PComm) when every rank can map
        # its peers' buffers (one node, <= 8 GPUs); otherwise ONE NCCL all-reduce of the flat buffer between two graphs
        self.p2p = _fused.p2p_enabled() and self.fused and 2 <= self.world <= 8 and dist.get_backend() == "nccl"
        self._p2p_comm = None
        with torch.no_grad():
            for t in self.model.state_dict().values():
                dist.broadcast(t, src=0)

    def p2p_comm(self, grad_floats, dev):
        """This rank's peer-mapped gradient buffer (None -- and the NCCL path from then on -- if IPC mapping fails)."""
        if self._p2p_comm is None and self.p2p:
            try:
                self._p2p_comm = _fused.P2PComm(grad_floats, dev, self.rank, self.world)
            except Exception as e:      # every rank raises together (see P2PComm)
                import warnings
                warnings.warn(f"gwnet_b200: {e}; using the NCCL all-reduce path", RuntimeWarning)
                self.p2p = False
        return self._p2p_comm

    def _allreduce_grads(self):
        import torch.distributed as dist
        flat = getattr(self.model, "_last_grad_flat", None)
        if getattr(self.model, "_flat", None) is not None:
            flat = self.model._flat.grad
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        lo, hi = (flat.data_ptr(), flat.data_ptr() + flat.numel() * 4) if flat is not None else (0, 0)
        if flat is not None and all(lo <= g.data_ptr() < hi for g in grads):
            dist.all_reduce(flat)                 # every p.grad is a view of this buffer
            flat.mul_(1.0 / self.world)
        else:
            for g in grads:
                dist.all_reduce(g)
                g.mul_(1.0 / self.world)

    def set_state(self, state):
        assert state == 'train' or state == 'val' or state == 'test'
        self.state = state

    # ---- the fork's synthetic multi-graph task (engine.py:64-117,132-180): the network predicts the fine signal, from
    # which a time-pooled signal F (mean over windows of F_t steps) and a cluster-pooled signal E (mean over the node
    # clusters of the sample's graph) are derived and compared with `real`.  The network runs on the native plan through
    # the autograd node of gwnet / gwnet_diff_G; the pooling is a handful of index ops on the [B,1,N,T] prediction.
    def _syn_forward(self, input, adj_idx):
        input = nn.functional.pad(input, (1, 0, 0, 0))
        if adj_idx is None:
            return self.model(input)
        assert self.state is not None, 'set train/val/test state first'
        supports = [s[adj_idx] for s in self.supports[self.state]]
        aptinit = self.aptinit[self.state] if self.aptinit is not None else None
        if aptinit is not None:
            aptinit = aptinit[adj_idx]
        return self.model(input, supports, aptinit)

    @staticmethod
    def _syn_pool(predict, F_t, G, adj_idx, pooltype):
        if pooltype != 'avg':
            raise NotImplementedError("pooltype 'subsample' is a TODO in the reference (engine.py:108-109)")
        lead = predict.shape[:-1]
        # F: window means over time, repeated back to full length
        F = predict.reshape(*lead, -1, F_t).mean(-1, keepdim=True).expand(*lead, predict.shape[-1] // F_t, F_t).reshape(*lead, -1)
        # E: cluster means over nodes, written back to the member nodes (clusters in dictionary order, like the reference)
        per_sample = type(G) == list
        parts = []
        for sample in (range(len(predict)) if per_sample else [None]):
            assign = (G[adj_idx[sample]] if per_sample else G).assign_dict
            e = predict[sample:sample + 1] if per_sample else predict
            for k in range(len(assign)):
                idx = torch.as_tensor(assign[k], dtype=torch.long, device=predict.device)
                e = e.index_copy(2, idx, e.index_select(2, idx).mean(2, keepdim=True).expand(-1, -1, idx.numel(), -1))
            parts.append(e)
        E = torch.cat(parts, 0) if per_sample else parts[0]
        return F, E

    def train_syn(self, input, real, F_t, G, adj_idx=None, pooltype='avg'):
        self.model.train()
        if isinstance(self.optimizer, _fused.FusedAdam):
            self._bind_flat(input)
            self.optimizer.zero_grad(set_to_none=False)
            self.optimizer.max_norm, self.optimizer.grad_scale = 0.0, 1.0
        else:
            self.optimizer.zero_grad()
        output = self._syn_forward(input, adj_idx).transpose(1, 3)
        predict = self.scaler.inverse_transform(output)
        F, predict = self._syn_pool(predict, F_t, G, adj_idx, pooltype)
        loss = self.loss(torch.cat((F, predict), 1), real, 0.0)
        loss.backward()
        if self.world > 1:
            self._allreduce_grads()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.optimizer.step()
        mape = util.masked_mape(predict, real, 0.0).item()
        rmse = util.masked_rmse(predict, real, 0.0).item()
        return loss.item(), mape, rmse

    def eval_syn(self, input, real, F_t, G, adj_idx=None, pooltype='avg'):
        assert type(G) != list or adj_idx is not None, 'adj index needed.'
        self.model.eval()
        with torch.no_grad():
            output = self._syn_forward(input, adj_idx if type(G) == list else None).transpose(1, 3)
            predict = self.scaler.inverse_transform(output)
            F, predict = self._syn_pool(predict, F_t, G, adj_idx, pooltype)
            loss = self.loss(torch.cat((F, predict), 1), real, 0.0)
            mape = util.masked_mape(predict, real, 0.0).item()
            rmse = util.masked_rmse(predict, real, 0.0).item()
        return loss.item(), mape, rmse, F, predict

    def _bind_flat(self, input):
        m = self.model
        if m._flat is None or not m._flat.intact():
            # the plan the forward below will use: the input is padded by one column first (engine.py:44)
            m._flat = _fused.FlatParams(m, m._runner(input.shape[0], input.shape[3] + 1).plan)
            self.optimizer._flat = None
        if not self.optimizer.bound:
            self.optimizer.bind(m._flat, int(torch.randint(0, 2 ** 62, (1,)).item()))

    def _fused_step(self, input, real_val):
        # scaler mean/std are baked into the captured launches by value: they are part of the key
        key = (tuple(input.shape), tuple(real_val.shape), self.model.precision, bool(self.use_graph), float(self.model.dropout),
               float(self.scaler.mean), float(self.scaler.std))
        st = self._steps.get(key)
        if st is None or not st.valid():
            self._steps = {k: v for k, v in self._steps.items() if v.valid()}
            st = _fused.FusedStep(self, input, real_val, self.use_graph)
            for v in self._steps.values():      # a new FlatParams re-homes the parameters: captured eval graphs are stale
                if isinstance(v, _fused.FusedEval) and not v.valid():
                    v.key = None
            self._steps[key] = st
        return st

    def train(self, input, real_val):
        self.model.train()
        if self.fused and self.loss is util.masked_mae:
            # forward + masked-MAE loss + backward + [all-reduce] + clip + Adam + metrics: one CUDA graph, one host sync.
            # (the +1 left pad of engine.py:44 and the receptive-field pad of model.py:176-180 are the same zero column,
            # folded into the start conv)
            return self._fused_step(input, real_val).run(input, real_val)
        # autograd path: a custom self.loss, or GWNET_B200_FUSED_STEP=0
        if isinstance(self.optimizer, _fused.FusedAdam):
            self._bind_flat(input)
            self.optimizer.zero_grad(set_to_none=False)     # p.grad are views of the flat gradient buffer: keep them
            self.optimizer.max_norm, self.optimizer.grad_scale = 0.0, 1.0
        else:
            self.optimizer.zero_grad()
        input = nn.functional.pad(input, (1, 0, 0, 0))
        output = self.model(input)
        output = output.transpose(1, 3)
        real = torch.unsqueeze(real_val, dim=1)
        predict = self.scaler.inverse_transform(output)
        loss = self.loss(predict, real, 0.0)
        loss.backward()
        if self.world > 1:
            self._allreduce_grads()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.optimizer.step()
        mape = util.masked_mape(predict, real, 0.0).item()
        rmse = util.masked_rmse(predict, real, 0.0).item()
        return loss.item(), mape, rmse

    def _fused_eval(self, input, real_val):
        key = ("eval", tuple(input.shape), tuple(real_val.shape), self.model.precision, bool(self.use_graph),
               float(self.scaler.mean), float(self.scaler.std))
        ev = self._steps.get(key)
        if ev is None or not ev.valid():
            self.model._table()
            ws = None
            for v in self._steps.values():      # share the forward workspace of a train step of the same shape
                if isinstance(v, _fused.FusedStep) and v.plan is self.model._runner(input.shape[0], input.shape[3] + 1).plan:
                    ws = v.workspace
            ev = _fused.FusedEval(self, input, real_val, self.use_graph, ws)
            self._steps[key] = ev
        return ev

    def eval(self, input, real_val):
        self.model.eval()
        if self.fused and self.loss is util.masked_mae:
            return self._fused_eval(input, real_val).run(input, real_val)
        input = nn.functional.pad(input, (1, 0, 0, 0))
        with torch.no_grad():   # the reference builds and discards a graph here; results are identical
            output = self.model(input)
        output = output.transpose(1, 3)
        real = torch.unsqueeze(real_val, dim=1)
        predict = self.scaler.inverse_transform(output)
        loss = self.loss(predict, real, 0.0)
        mape = util.masked_mape(predict, real, 0.0).item()
        rmse = util.masked_rmse(predict, real, 0.0).item()
        return loss.item(), mape, rmse
