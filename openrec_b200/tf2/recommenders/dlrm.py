"""DLRM -- mirrors openrec/tf2/recommenders/dlrm.py:6-100 on liborx (gathers, interaction, Dense layers,
loss, sparse + dense optimizer applies); same lazy step protocol as the other recommenders."""
import sys

import torch

from ... import native as N
from ...tfshim.core import LazyScalar, StepNode, Tensor, convert
from ...tfshim.keras import Model
from ..mlp_ops import ACT, DLRMGraph
from ..modules import MLP, LatentFactor, SecondOrderFeatureInteraction


class DLRM(Model):
    def __init__(self, m_spa, ln_emb, ln_bot, ln_top, arch_interaction_op="dot", arch_interaction_itself=False,
                 sigmoid_bot=False, sigmoid_top=True, loss_func="mse", loss_threshold=0.0,
                 interaction_mode="reference"):
        """Reference signature (dlrm.py:8-19) + ``interaction_mode``: 'reference' reproduces the reference's
        dot interaction bit for bit (identically zero), 'dlrm' is the strictly-lower triangle."""
        super().__init__()
        self._m_spa = int(m_spa)
        self._loss_threshold = loss_threshold
        self._loss_func = loss_func
        self._latent_factors = [LatentFactor(num_instances=int(num), dim=m_spa) for num in ln_emb]
        self._mlp_bot = MLP(units_list=ln_bot, out_activation="sigmoid" if sigmoid_bot else "relu")
        self._mlp_top = MLP(units_list=ln_top, out_activation="sigmoid" if sigmoid_top else "relu")
        self._dot_interaction = None
        if arch_interaction_op == "dot":
            self._dot_interaction = SecondOrderFeatureInteraction(self_interaction=arch_interaction_itself,
                                                                  mode=interaction_mode)
        elif self._arch_interaction_op != "cat":   # the reference never assigns this attribute: AttributeError (Q2)
            sys.exit("ERROR: arch_interaction_op=" + self._arch_interaction_op + " is not supported")
        if loss_func not in ("mse", "bce"):
            sys.exit("ERROR: loss_func=" + loss_func + " is not supported")
        self._self_interaction = bool(arch_interaction_itself)
        self._interaction_mode = interaction_mode

    # ---- graph over the current variables
    def _graph(self, n_dense):
        T = len(self._latent_factors)
        width = T + 1
        P = width * (width + 1) // 2 if self._self_interaction else width * (width - 1) // 2
        self._mlp_bot.build(n_dense)
        self._mlp_top.build(self._m_spa + P)

        def layers(mlp):
            return [(l.kernel.t, None if l.bias is None else l.bias.t, ACT[l.activation]) for l in mlp.layers]
        clip = float(self._loss_threshold) if 0.0 < self._loss_threshold < 1.0 else 0.0
        return DLRMGraph([lf.embeddings.t for lf in self._latent_factors], layers(self._mlp_bot),
                         layers(self._mlp_top), self._m_spa, self._self_interaction, self._interaction_mode,
                         0 if self._loss_func == "mse" else 1, clip)

    @staticmethod
    def _inputs(dense_features, sparse_features, label=None):
        dense = convert(dense_features).t.to(torch.float32).contiguous()
        sparse = convert(sparse_features).t.to(torch.int32).contiguous()   # Embedding casts ids to int32
        lab = None if label is None else convert(label).t.to(torch.float32).reshape(-1).contiguous()
        return dense, sparse, lab

    def call(self, dense_features, sparse_features, label):
        """-> loss (a single lazy scalar, dlrm.py:63-74)."""
        node = StepNode(self, 1)
        node.inputs = self._inputs(dense_features, sparse_features, label)
        self._graph(node.inputs[0].shape[1])   # keras builds the Dense layers during the first call
        return LazyScalar(node, {0: 1.0})

    def inference(self, dense_features, sparse_features):
        """-> predictions [B] (dlrm.py:76-100)."""
        dense, sparse, _ = self._inputs(dense_features, sparse_features)
        return Tensor(self._graph(dense.shape[1]).forward(dense, sparse)["pred"])

    # ---- step protocol hooks (openrec_b200/tfshim/core.py)
    def _orx_step_variables(self):
        return self.trainable_variables

    def _orx_forward(self, node):
        dense, sparse, lab = node.inputs
        node.out.copy_(self._graph(dense.shape[1]).forward(dense, sparse, lab)["out4"])

    def _fwd_bwd(self, node, scale):
        dense, sparse, lab = node.inputs
        g = self._graph(dense.shape[1])
        c = g.forward(dense, sparse, lab, want_grad=True)
        if scale != 1.0:
            c["dpred"].mul_(scale)
        return c, g.backward(c)

    def _orx_apply(self, node, grads_and_vars, optimizer):
        if node.stepped:
            raise RuntimeError("this model call's gradients were already applied")
        dense, sparse, lab = node.inputs
        self._graph(dense.shape[1])                                   # make sure Dense layers exist
        want = {id(v) for v in self.trainable_variables}
        got = {id(v) for _, v in grads_and_vars}
        coefs = [g.coef for g, _ in grads_and_vars]
        if got != want or any(c != coefs[0] for c in coefs):
            raise NotImplementedError("apply_gradients: DLRM's fused step needs the gradients of ALL trainable "
                                      "variables w.r.t. one objective")
        c, (dZ, bot_g, top_g) = self._fwd_bwd(node, float(coefs[0].get(0, 0.0)))
        eng, o = N.engine(), optimizer.opt_struct()
        for k, lf in enumerate(self._latent_factors):                 # IndexedSlices(ids = sparse[:,k], dZ[:,k,:])
            eng.sparse_apply_strided(optimizer.table(lf.embeddings), sparse, k, dZ, o)
        for mlp, grads in ((self._mlp_bot, bot_g), (self._mlp_top, top_g)):
            for layer, (dw, db) in zip(mlp.layers, grads):
                eng.dense_apply(layer.kernel.t, *optimizer.slots(layer.kernel), dw, o)
                if layer.bias is not None:
                    eng.dense_apply(layer.bias.t, *optimizer.slots(layer.bias), db, o)
        node.out = c["out4"]
        node.stepped = True
        node.inputs = None

    def _launches_per_step(self):
        """liborx kernel launches of one training step (bench.py's gpu_launches): per table gather + index / apply / tail,
        per Dense layer forward (1) + backward (activation, column sum, dgrad, wgrad; split-K adds a reduce) + 2 dense
        applies, 2 interaction kernels, 1 loss kernel."""
        T = len(self._latent_factors)
        n_dense = len(self._mlp_bot.layers) + len(self._mlp_top.layers)
        return 4 * T + n_dense * (1 + 4 + 2) + 2 + 1

    def _orx_materialize_grad(self, node, var, coef):
        if node.stepped:
            raise RuntimeError("gradients requested after the step was applied")
        _, (dZ, bot_g, top_g) = self._fwd_bwd(node, float(coef.get(0, 0.0)))
        for k, lf in enumerate(self._latent_factors):
            if var is lf.embeddings:
                return Tensor(node.inputs[1][:, k].contiguous()), Tensor(dZ[:, k, :].contiguous())
        for mlp, grads in ((self._mlp_bot, bot_g), (self._mlp_top, top_g)):
            for layer, (dw, db) in zip(mlp.layers, grads):
                if var is layer.kernel:
                    return None, Tensor(dw)
                if var is layer.bias:
                    return None, Tensor(db)
        raise KeyError("variable does not belong to this model")
