"""Drop-in replacement for the reference's `fusions` package (same module and
class names, constructor signatures, parameter names and error behaviour --
reference: /fusions/*.py), backed by the immtsf sm_100a kernels."""
