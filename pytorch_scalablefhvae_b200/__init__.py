"""B200-native ScalableFHVAE train / inference step (drop-in for BurnhamG/PyTorch-ScalableFHVAE's
``SimpleFHVAE`` / ``FHVAE`` modules).  Hand-written sm_100a CUDA behind a C ABI; no CPU fallback."""
from . import _lib
from ._lib import MODE_BF16, MODE_BF16X3, MODE_F32_SIMT, build, set_deterministic
from .model import FHVAE, SimpleFHVAE, loss_function
from .optim import FusedAdam
from .inference import extract_posteriors, extract_posteriors_sharded, segment_table, shard_utterances
from .hierarchical import HierarchicalTrainer, ShardedMu2Table, sample_sequences
from .parallel import DataParallel, shard_alloc_rows
from .checkpoint import load_checkpoint_file, save_checkpoint

__all__ = ["FHVAE", "SimpleFHVAE", "FusedAdam", "loss_function", "build", "MODE_F32_SIMT", "MODE_BF16X3",
           "MODE_BF16", "extract_posteriors", "segment_table", "ShardedMu2Table", "sample_sequences",
           "load_checkpoint_file", "save_checkpoint", "HierarchicalTrainer", "DataParallel", "shard_alloc_rows",
           "extract_posteriors_sharded", "shard_utterances", "set_deterministic"]
