"""cnvetti_b200 -- B200-native (sm_100a CUDA) read-depth hot path of bihealth/cnvetti.

The package holds only what the hot path needs: `csrc/` (hand-written kernels + the C ABI of
include/cnvetti_b200.h) and this thin host-side mirror of the reference's operator interface.
"""
from ._native import CnvbError, ReferencePanic  # noqa: F401
from .context import Context, default_context  # noqa: F401
from .coverage import (AggregationStats, BamRecordAggregator, CoverageAggregator,  # noqa: F401
                       CoverageOptions, FragmentsGenomeWideAggregator,
                       FragmentsTargetRegionsAggregator, ReadBatch, ReferenceStats, cov_records,
                       make_aggregator, pack_reads)
from . import discover, model_wis, piles  # noqa: F401
from .model_wis import BuildModelWisOptions  # noqa: F401
from .normalize import run_binned_correction, run_uncorrected_normalization  # noqa: F401

__version__ = "0.1.0"
