from unicore.models import register_model, register_model_architecture
from unicore_b200.models.unimol import UniMolModel as _Impl
from unicore_b200.models.unimol import apply_unimol_arch


@register_model("unimol")
class UniMolModel(_Impl):
    pass


@register_model_architecture("unimol", "unimol")
def base_architecture(args):
    apply_unimol_arch(args)


@register_model_architecture("unimol", "unimol_base")
def unimol_base_architecture(args):
    apply_unimol_arch(args)
