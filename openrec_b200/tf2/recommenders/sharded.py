"""ShardedBPR / ShardedUCML -- the BPR / UCML class surface (openrec/tf2/recommenders/bpr.py:7-19, ucml.py:7-19: same
constructor arguments, same ``model(user_id, p_item_id, n_item_id) -> (loss, l2_loss)``, same GradientTape /
``optimizer.apply_gradients`` step protocol) on ROW-SHARDED tables: one process per GPU under ``torch.distributed``,
row r of the user / item tables on rank ``r % world_size`` (openrec_b200/sharded.py, csrc/orx_shard.cu).  The reference is
single-device; this is what lets the 100M-item table of BASELINE configs[4] exist at all.

Every rank calls the model with ITS part of the global batch (any user / item ids); ``apply_gradients`` runs the one
sharded step; the (loss, l2_loss) it returns are those of the GLOBAL batch, identical on every rank.  The model's
variables are the local shards; the keras optimizer owns the slot tensors, so ``openrec_b200.tf2.checkpoint`` saves and
restores a rank's shard like any other model (one file per rank).  SGD, Adagrad and LazyAdam (row-sparse Adam; Keras'
``Adam()`` sweeps whole tables every step and is not offered sharded)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ... import native as N
from ...sharded import HomeRoutedPairwise
from ...tfshim.core import LazyScalar, StepNode, Variable
from ...tfshim.keras import Model
from ._base import ids_of


class _Shard:
    """Stand-in for the LatentFactor attribute of the reference models: ``.embeddings`` / ``.variables[0]`` is this rank's
    shard ([rows r with r % world == rank, dim])."""

    def __init__(self, var, total, dim):
        self.embeddings, self.input_dim, self.output_dim = var, total, dim
        self.variables = self.trainable_variables = [var]


class ShardedBPR(Model):
    _kind = N.ORX_PAIR_BPR

    def __init__(self, dim_user_embed, dim_item_embed, total_users, total_items, seed=0):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("ShardedBPR needs torch.distributed (one process per GPU; world size 1 is allowed)")
        if dim_user_embed != dim_item_embed:
            raise ValueError("user and item embedding dims must match (the reference multiplies them elementwise)")
        self._rank, self._world = dist.get_rank(), dist.get_world_size()
        self._U, self._I, self._D = int(total_users), int(total_items), int(dim_user_embed)
        eng = self._eng = N.engine()
        r, R = self._rank, self._world
        ru, ri = (self._U - r + R - 1) // R, (self._I - r + R - 1) // R
        mk = lambda rows, cols, k, name: self._new_var(eng, rows, cols, seed * 1000003 + r * 17 + k, name)
        self.user_latent_factor = _Shard(mk(max(ru, 1), self._D, 0, "user_latent_factor"), self._U, self._D)
        self.item_latent_factor = _Shard(mk(max(ri, 1), self._D, 1, "item_latent_factor"), self._I, self._D)
        self.item_bias = _Shard(mk(max(ri, 1), 1, 2, "item_bias"), self._I, 1)
        self._impl = None
        self._impl_key = None

    @staticmethod
    def _new_var(eng, rows, cols, seed, name):
        t = torch.empty(rows, cols, dtype=torch.float32, device=eng.device)
        eng.fill_uniform(t, -0.05, 0.05, seed)                       # LatentFactor's 'uniform' initializer
        v = Variable.__new__(Variable)
        v.t, v.trainable, v.name = t, True, name
        return v

    def _get_margin(self):
        return 0.0

    @property
    def variables(self):
        return [self.user_latent_factor.embeddings, self.item_latent_factor.embeddings, self.item_bias.embeddings]

    trainable_variables = variables

    def _orx_step_variables(self):
        return self.variables

    def call(self, user_id, p_item_id, n_item_id):
        """-> (loss, l2_loss) of the GLOBAL batch as lazy scalars; this rank contributes the triplets it was given."""
        node = StepNode(self, 2)
        node.ids = tuple(ids_of(x) for x in (user_id, p_item_id, n_item_id))
        return LazyScalar(node, {0: 1.0}), LazyScalar(node, {1: 1.0})

    def _orx_forward(self, node):
        raise NotImplementedError("a sharded model's loss exists only as part of the training step "
                                  "(read it after optimizer.apply_gradients)")

    def _orx_materialize_grad(self, node, var, coef):
        raise NotImplementedError("explicit IndexedSlices are not available for row-sharded tables")

    def _orx_apply(self, node, grads_and_vars, optimizer):
        if node.stepped:
            raise RuntimeError("this model call's gradients were already applied")
        want = {id(v) for v in self.variables}
        coefs = [g.coef for g, _ in grads_and_vars]
        if {id(v) for _, v in grads_and_vars} != want or any(c != coefs[0] for c in coefs):
            raise NotImplementedError("apply_gradients: the sharded step needs the gradients of ALL of the model's "
                                      "variables w.r.t. one objective")
        kind = optimizer._kind
        if kind not in (N.ORX_OPT_SGD, N.ORX_OPT_ADAGRAD, N.ORX_OPT_ADAM_LAZY):
            raise NotImplementedError("sharded tables: use SGD, Adagrad or LazyAdam (Keras Adam() sweeps whole tables)")
        B = node.ids[0].numel()
        key = (id(optimizer), kind)
        if self._impl is None or self._impl_key != key or B > self._impl.B:
            if self._impl is not None:
                self._impl.close()
            vs = self.variables
            self._impl = HomeRoutedPairwise(self._eng, self._rank, self._world, self._U, self._I, self._D, B, kind=self._kind,
                                            opt_kind=kind, tables=tuple(v.t for v in vs),
                                            slots=tuple(optimizer.slots(v) for v in vs))
            self._impl_key, self._impl_opt = key, optimizer          # strong reference: the slots' raw pointers are in use
        m = self._impl
        m.lr, m.eps, m.b1, m.b2 = optimizer.learning_rate, optimizer.epsilon, optimizer.beta_1, optimizer.beta_2
        m.margin = self._get_margin()
        m.iterations = optimizer.iterations - 1                      # HomeRoutedPairwise.step increments it
        out = m.step(*node.ids, c_loss=float(coefs[0].get(0, 0.0)), c_l2=float(coefs[0].get(1, 0.0)))
        node.out = m._out[m.iterations % 16]
        node.stepped = True
        node.ids = None

    def inference(self, user_id):
        raise NotImplementedError("full-catalogue scoring needs the whole item table on one device")

    def check(self):
        """Raise if the sharded step flagged an error (peer timeout, mailbox overflow)."""
        if self._impl is not None:
            self._impl.check()


class ShardedUCML(ShardedBPR):
    _kind = N.ORX_PAIR_UCML

    def __init__(self, dim_user_embed, dim_item_embed, total_users, total_items, margin=0.5, seed=0):
        super().__init__(dim_user_embed, dim_item_embed, total_users, total_items, seed=seed)
        self.margin = margin

    def _get_margin(self):
        return float(self.margin)
