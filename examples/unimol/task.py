"""Synthetic molecules for the Uni-Mol style benchmark: random atom types and 3-D conformers, 15 % of
the atoms masked with Gaussian coordinate noise, pairwise distances and edge types derived on the fly,
2-D padded collation (multiple of 8)."""
import numpy as np
import torch

from unicore.data import Dictionary, UnicoreDataset, data_utils
from unicore.tasks import UnicoreTask, register_task


class SyntheticMoleculeDataset(UnicoreDataset):
    def __init__(self, n, dictionary, mask_idx, min_atoms, max_atoms, seed, mask_prob=0.15, noise=1.0):
        self.n, self.d, self.mask_idx = n, dictionary, mask_idx
        self.min_atoms, self.max_atoms, self.seed = min_atoms, max_atoms, seed
        self.mask_prob, self.noise = mask_prob, noise
        self.epoch = 1
        special = set(dictionary.special_index()) | {mask_idx}
        self.atoms = np.array([i for i in range(len(dictionary)) if i not in special])

    def set_epoch(self, epoch):
        self.epoch = epoch

    @property
    def can_reuse_epoch_itr_across_epochs(self):
        return True

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        rng = np.random.RandomState((self.seed * 7919 + self.epoch * 104729 + i) % (2 ** 31 - 1))
        n = rng.randint(self.min_atoms, self.max_atoms + 1)
        tokens = np.concatenate([[self.d.bos()], self.atoms[rng.randint(0, len(self.atoms), n)], [self.d.eos()]])
        coord = np.concatenate([np.zeros((1, 3)), rng.randn(n, 3) * 3.0, np.zeros((1, 3))]).astype(np.float32)
        n_mask = max(1, int(round(self.mask_prob * n)))
        pos = rng.choice(n, n_mask, replace=False) + 1
        target = np.full(len(tokens), self.d.pad(), dtype=np.int64)
        target[pos] = tokens[pos]
        src = tokens.copy()
        src[pos] = self.mask_idx
        noisy = coord.copy()
        noisy[pos] += rng.randn(n_mask, 3).astype(np.float32) * self.noise
        dist = np.linalg.norm(noisy[:, None] - noisy[None], axis=-1).astype(np.float32)
        dist_t = np.linalg.norm(coord[:, None] - coord[None], axis=-1).astype(np.float32)
        edge = src[:, None] * len(self.d) + src[None, :]
        return {
            "src_tokens": torch.from_numpy(src.astype(np.int64)), "src_coord": torch.from_numpy(noisy),
            "src_distance": torch.from_numpy(dist), "src_edge_type": torch.from_numpy(edge.astype(np.int64)),
            "tokens_target": torch.from_numpy(target), "coord_target": torch.from_numpy(coord),
            "distance_target": torch.from_numpy(dist_t),
        }

    def collater(self, samples):
        if len(samples) == 0:
            return {}
        pad = self.d.pad()
        tok = data_utils.collate_tokens([s["src_tokens"] for s in samples], pad, pad_to_multiple=8)
        L = tok.size(1)

        def pad1(key, value):
            return data_utils.collate_tokens([s[key] for s in samples], value, pad_to_multiple=8)

        def pad_coord(key):
            out = samples[0][key].new_zeros(len(samples), L, 3)
            for i, s in enumerate(samples):
                out[i, : s[key].size(0)] = s[key]
            return out

        def pad2(key, value):
            return data_utils.collate_tokens_2d([s[key] for s in samples], value, pad_to_multiple=8)

        return {
            "net_input": {"src_tokens": tok, "src_coord": pad_coord("src_coord"),
                          "src_distance": pad2("src_distance", 0), "src_edge_type": pad2("src_edge_type", 0)},
            "target": {"tokens_target": pad1("tokens_target", pad), "coord_target": pad_coord("coord_target"),
                       "distance_target": pad2("distance_target", 0)},
        }


@register_task("synthetic_unimol")
class SyntheticUniMolTask(UnicoreTask):
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
        n = self.args.synthetic_num_samples if split == self.args.train_subset else max(8, self.args.synthetic_num_samples // 8)
        self.datasets[split] = SyntheticMoleculeDataset(
            n, self.dictionary, self.mask_idx, self.args.synthetic_min_atoms, self.args.synthetic_max_atoms,
            seed=self.args.seed + (0 if split == self.args.train_subset else 1), mask_prob=self.args.mask_prob,
            noise=self.args.noise,
        )
