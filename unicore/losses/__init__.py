"""Loss registry (``--loss``, default ``cross_entropy``). Reference ``unicore/losses/__init__.py``."""
from unicore import registry
from unicore.losses.unicore_loss import UnicoreLoss

build_loss_, register_loss, LOSS_REGISTRY = registry.setup_registry(
    "--loss", base_class=UnicoreLoss, default="cross_entropy"
)


def build_loss(args, task):
    return build_loss_(args, task)


from . import cross_entropy, masked_lm  # noqa: E402,F401
