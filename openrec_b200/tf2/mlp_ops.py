"""Dense / interaction operators behind keras Dense, MLP and SecondOrderFeatureInteraction (DLRM path).
Filled in by the DLRM kernels; until then these fail loudly (no CPU / torch fallback)."""


def dense_forward(x, kernel, bias, activation):
    raise NotImplementedError("DLRM MLP kernels are not built yet")


def interaction_forward(feats, self_interaction, mode):
    raise NotImplementedError("DLRM interaction kernel is not built yet")
