"""Synthetic models and index-addressable data streams for tests, golden fixtures and bench.py.

Not part of the product: random-init modules of the shapes BASELINE.json names (there is no
network for checkpoints or datasets) and seeded streams following SURVEY.md section 8(d).
"""
from .models import (ConvMLPNet, DeiTLike, LlamaLikeDecoder, PrimitiveConv1x1Net,  # noqa: F401
                     PrimitiveLinearNet, llama_ce_loss)
from .streams import (IndexedStream, image_batch, lowrank_image_batch, step_spectrum_activations,  # noqa: F401
                      token_batch)
