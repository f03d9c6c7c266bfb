"""Import stub (test infrastructure): yastn/__init__.py imports opt_einsum for its optional path-search module
(yastn/tensor/oe_blocksparse.py:23-24), which nothing on the contraction hot path uses and which is not installed here."""
from . import contract  # noqa: F401
