"""Synthetic-molecule task of the portable plug-in: the dataset of ``examples/unimol/task.py`` (public data API only)
under its own task name, so that both plug-ins can be loaded side by side."""
import importlib.util
import os

from unicore.data import Dictionary
from unicore.tasks import UnicoreTask, register_task

_spec = importlib.util.spec_from_file_location(
    "_unimol_synthetic_data", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "unimol", "task.py"))


def _dataset_class():
    # the dataset class lives next door; importing that file also registers its own task there, which is harmless
    # under this framework and impossible under the reference (duplicate names are refused) - so copy the class only
    src = open(_spec.origin).read().split("@register_task")[0]
    scope = {}
    exec(compile(src, _spec.origin, "exec"), scope)  # noqa: S102
    return scope["SyntheticMoleculeDataset"]


SyntheticMoleculeDataset = _dataset_class()


@register_task("synthetic_unimol_portable")
class PortableSyntheticUniMolTask(UnicoreTask):
    @staticmethod
    def add_args(parser):
        parser.add_argument("data", nargs="?", default=None)
        parser.add_argument("--synthetic-num-samples", default=4096, type=int)
        parser.add_argument("--synthetic-atom-types", default=30, type=int)
        parser.add_argument("--synthetic-min-atoms", default=64, type=int)
        parser.add_argument("--synthetic-max-atoms", default=254, type=int)
        parser.add_argument("--mask-prob", default=0.15, type=float)
        parser.add_argument("--noise", default=1.0, type=float)

    def __init__(self, args, dictionary):
        super().__init__(args)
        self.dictionary = dictionary
        self.seed = args.seed
        self.mask_idx = dictionary.add_symbol("[MASK]", is_special=True)

    @classmethod
    def setup_task(cls, args, **kwargs):
        d = Dictionary()
        for s in ("[PAD]", "[CLS]", "[SEP]", "[UNK]"):
            d.add_symbol(s, is_special=True)
        for i in range(args.synthetic_atom_types):
            d.add_symbol("A{}".format(i))
        return cls(args, d)

    def load_dataset(self, split, **kwargs):
        a = self.args
        train = split == a.train_subset
        n = a.synthetic_num_samples if train else max(8, a.synthetic_num_samples // 8)
        self.datasets[split] = SyntheticMoleculeDataset(
            n, self.dictionary, self.mask_idx, a.synthetic_min_atoms, a.synthetic_max_atoms,
            seed=a.seed + (0 if train else 1), mask_prob=a.mask_prob, noise=a.noise)
