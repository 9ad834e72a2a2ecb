"""Flagship models built on the fused op set: BERT (masked LM) and a Uni-Mol style SE(3)
transformer.  They are plain ``BaseUnicoreModel`` subclasses; registration under CLI names happens
in the plug-in packages under ``examples/`` (mirroring how the reference ships its BERT)."""
from .bert import BertModel, BertLMHead, BertClassificationHead, apply_arch  # noqa: F401
