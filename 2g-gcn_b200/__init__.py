"""2g-gcn_b200 — B200-native (sm_100a) implementation of the 2G-GCN forward/backward hot path.

The directory name is the one the build contract prescribes; because it is not a valid Python
identifier, import it with ``importlib.import_module('2g-gcn_b200')`` or through the alias module
``tggcn_b200`` at the repository root.

Public surface (mirrors the reference's ``vhoi.models`` for this path):
    TGGCN, select_model, install_dropin      -- drop-in model class (model.py)
    abi                                      -- ctypes binding of include/tggcn_b200.h
    synth                                    -- synthetic MPHOI/CAD-120/Bimanual-shaped batches
    dp                                       -- data-parallel glue: bucketed gradient all-reduce overlapped with the backward, valid-count loss weights
    losses                                   -- fused criterion, drop-in for vhoi.losses.select_loss (budget / BCE / NLL in two kernels)
    feeder                                   -- device-resident dataset; double-buffered host->device input pipeline (pinned memory, side stream)
    trainer                                  -- data-parallel training driver (epoch loop, sharded sampler, checkpoint dict of train_utils.train)
    evaluate                                 -- device-side predict.py post-processing: up-sampling + argmax, F1@k (pyrutils/metrics.py)
    build                                    -- in-tree nvcc build of lib2ggcn_b200.so
"""
from . import abi, dp, evaluate, feeder, losses, optim, synth, trainer        # noqa: F401
from .model import TGGCN, select_model, install_dropin   # noqa: F401

__all__ = ['TGGCN', 'select_model', 'install_dropin', 'abi', 'dp', 'evaluate', 'feeder', 'losses', 'optim', 'synth', 'trainer']
