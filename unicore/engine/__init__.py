"""The training engine behind ``unicore.trainer.Trainer``.

``Trainer`` is a thin façade with the reference's public names (``unicore/trainer.py:30`` there); the work is done by

* ``replica.Replica``    - the model / loss pair: precision, device, re-tied shared parameters, data-parallel wrapping
* ``ledger.StatLedger``  - one update's logging statistics as a flat fp64 vector and its cross-rank sum (inside the fused
                           optimizer tail, on the peer-memory kernel, or through the process group)
* ``update.UpdateStep``  - one optimizer update as explicit phases over device-resident step state
* ``snapshot``           - checkpoint state assembly and restoration (reference schema)
* ``clock.TrainingClock``- wall-clock bookkeeping that survives restarts
"""
from .clock import TrainingClock  # noqa: F401
from .ledger import StatLedger  # noqa: F401
from .replica import Replica  # noqa: F401
from .update import LazyStats, UpdateStep  # noqa: F401
