"""Drop-in replacement for the reference's ``models`` package (same module and symbol names)."""
from . import positional_encoding, rope_utils, vit  # noqa: F401
from .positional_encoding import (AbsolutePositionalEncoding, NoPositionalEncoding, PolynomialRPE,  # noqa: F401
                                  RelativePositionalEncoding, RoPEAxial, RoPEMixed)
from .rope_utils import apply_rotary_emb, reshape_for_broadcast  # noqa: F401
from .vit import Attention, Block, VisionTransformer  # noqa: F401
