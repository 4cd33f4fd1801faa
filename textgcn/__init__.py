"""Drop-in import surface of the reference package (`textgcn/__init__.py`):
`from textgcn import Text2GraphTransformer` / `from textgcn import models`, backed by pytextgcn_b200."""
from .lib import Text2GraphTransformer
from .lib import models

__all__ = ["Text2GraphTransformer", "models"]
