"""Table-driven registration of the plain ``torch.optim`` optimizers (``sgd``, ``adagrad``, ``adadelta``).

The reference ships one hand-written wrapper class per optimizer (``unicore/optim/sgd.py:13``, ``adagrad.py:13``,
``adadelta.py:13``);
they only differ in the torch class, the flags they contribute and how flags map to constructor keywords.
Here that is data: ``_SPECS`` lists, per registry name, the torch class, its flags (``(dest, flag names,
argparse keywords)``), the constructor-keyword -> ``args`` attribute mapping and whether the update works on a
flattened parameter vector.  ``make_wrapper`` builds and registers the ``UnicoreOptimizer`` subclass.
"""
import torch.optim

from . import UnicoreOptimizer, register_optimizer

_WEIGHT_DECAY = ("weight_decay", ("--weight-decay", "--wd"), dict(default=0.0, type=float, metavar="WD", help="weight decay"))

_SPECS = {
    "sgd": dict(
        cls=torch.optim.SGD,
        class_name="SGD",
        flags=[("momentum", ("--momentum",), dict(default=0.0, type=float, metavar="M", help="momentum factor")),
               _WEIGHT_DECAY],
        kwargs={"momentum": "momentum", "weight_decay": "weight_decay"},
        flat=True,
    ),
    "adagrad": dict(
        cls=torch.optim.Adagrad,
        class_name="Adagrad",
        flags=[_WEIGHT_DECAY],
        kwargs={"weight_decay": "weight_decay"},
        flat=False,
    ),
    "adadelta": dict(
        cls=torch.optim.Adadelta,
        class_name="Adadelta",
        flags=[("adadelta_rho", ("--adadelta-rho",), dict(type=float, default=0.9, metavar="RHO",
                                                           help="decay of the running average of squared gradients")),
               ("adadelta_eps", ("--adadelta-eps",), dict(type=float, default=1e-6, metavar="EPS",
                                                           help="denominator term for numerical stability")),
               _WEIGHT_DECAY,
               ("anneal_eps", ("--anneal-eps",), dict(action="store_true", help="flag to anneal eps"))],
        kwargs={"rho": "adadelta_rho", "eps": "adadelta_eps", "weight_decay": "weight_decay"},
        flat=True,
    ),
}


def make_wrapper(name):
    spec = _SPECS[name]

    def __init__(self, args, params):
        UnicoreOptimizer.__init__(self, args)
        self._optimizer = spec["cls"](params, **self.optimizer_config)

    def add_args(parser):
        for _dest, names, kw in spec["flags"]:
            parser.add_argument(*names, **kw)

    def optimizer_config(self):
        cfg = {"lr": self.args.lr[0]}
        cfg.update({key: getattr(self.args, attr) for key, attr in spec["kwargs"].items()})
        return cfg

    namespace = {
        "__init__": __init__,
        "__doc__": "``torch.optim.{}`` behind the UnicoreOptimizer interface (``--optimizer {}``).".format(
            spec["cls"].__name__, name),
        "add_args": staticmethod(add_args),
        "optimizer_config": property(optimizer_config),
        "supports_flat_params": property(lambda self: spec["flat"]),
    }
    return register_optimizer(name)(type(spec["class_name"], (UnicoreOptimizer,), namespace))


SGD = make_wrapper("sgd")
Adagrad = make_wrapper("adagrad")
Adadelta = make_wrapper("adadelta")
