"""Stand-in for torch-topological 0.1.7 (see ../README.md): only what the reference's topological_loss.py imports."""
__version__ = "0.1.7-standin"
