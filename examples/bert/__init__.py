"""BERT plug-in for ``--user-dir examples/bert``: registers task ``bert`` and model ``bert`` with
architectures ``bert`` / ``bert_base`` / ``bert_large`` / ``xlm`` (counterpart of the reference's
``examples/bert``; the implementation lives in ``unicore_b200.models.bert``)."""
from . import task, model  # noqa: F401
