"""Masked-LM task over an LMDB of raw text lines (reference ``examples/bert/task.py:31-124``):
``<data>/dict.txt`` vocabulary, ``<data>/<split>.lmdb`` records, WordPiece tokenisation, 15 %
BERT masking, right-padded batches in a fixed random order."""
import logging
import os

import numpy as np

from unicore.data import (
    BertTokenizeDataset,
    Dictionary,
    LMDBDataset,
    MaskTokensDataset,
    NestedDictionaryDataset,
    RightPadDataset,
    SortDataset,
    data_utils,
)
from unicore.tasks import UnicoreTask, register_task

logger = logging.getLogger(__name__)


@register_task("bert")
class BertTask(UnicoreTask):
    @staticmethod
    def add_args(parser):
        parser.add_argument("data", help="directory with dict.txt and <split>.lmdb files")
        parser.add_argument("--mask-prob", default=0.15, type=float, help="probability of replacing a token with mask")
        parser.add_argument("--leave-unmasked-prob", default=0.1, type=float,
                            help="probability that a masked token is unmasked")
        parser.add_argument("--random-token-prob", default=0.1, type=float,
                            help="probability of replacing a token with a random token")

    def __init__(self, args, dictionary):
        super().__init__(args)
        self.dictionary = dictionary
        self.seed = args.seed
        self.mask_idx = dictionary.add_symbol("[MASK]", is_special=True)

    @classmethod
    def setup_task(cls, args, **kwargs):
        dictionary = Dictionary.load(os.path.join(args.data, "dict.txt"))
        logger.info("dictionary: {} types".format(len(dictionary)))
        return cls(args, dictionary)

    def load_dataset(self, split, combine=False, **kwargs):
        records = LMDBDataset(os.path.join(self.args.data, split + ".lmdb"))
        tokens = BertTokenizeDataset(records, os.path.join(self.args.data, "dict.txt"), max_seq_len=self.args.max_seq_len)
        src, tgt = MaskTokensDataset.apply_mask(
            tokens,
            self.dictionary,
            pad_idx=self.dictionary.pad(),
            mask_idx=self.mask_idx,
            seed=self.args.seed,
            mask_prob=self.args.mask_prob,
            leave_unmasked_prob=self.args.leave_unmasked_prob,
            random_token_prob=self.args.random_token_prob,
        )
        with data_utils.numpy_seed(self.args.seed):
            order = np.random.permutation(len(src))
        pad = self.dictionary.pad()
        batchable = NestedDictionaryDataset(
            {"net_input": {"src_tokens": RightPadDataset(src, pad_idx=pad)}, "target": RightPadDataset(tgt, pad_idx=pad)}
        )
        self.datasets[split] = SortDataset(batchable, sort_order=[order])

    def build_model(self, args):
        from unicore import models

        return models.build_model(args, self)
