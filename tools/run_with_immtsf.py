#!/usr/bin/env python
"""Run an UNMODIFIED IMM-TSF script (main.py, main_all.py) with the B200 fusion modules swapped in.

    python tools/run_with_immtsf.py /path/to/IMM-TSF/main.py --model tPatchGNN --TTF_module TTF_T2V_XAttn ...

How: the reference resolves the fusion modules with `from fusions.FusionModel import FusionModel`
(main.py:39) and `from fusions.load_llm import get_context_window_size` (main.py:40).  Python puts the script's
directory first on sys.path, so a same-named package elsewhere on the path would lose; but a module already in
`sys.modules` wins over any path entry.  This launcher imports the drop-in `fusions` package (and its
submodules) first, then executes the script with runpy.  argparse `choices` (main.py:620-633) are unchanged because
the drop-in registers under the reference's own four names.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "imm-tsf_b200")


def install():
    """Make `import fusions...` resolve to the B200 drop-in for the rest of this process.  Also switches on the transparent
    CUDA-graph replay of FusionModel.forward / backward (immtsf/autograph.py) unless IMMTSF_AUTOGRAPH is already set: the
    reference's loop then runs the fusion path at graph speed without any change to main.py."""
    os.environ.setdefault("IMMTSF_AUTOGRAPH", "1")
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    for name in [m for m in sys.modules if m == "fusions" or m.startswith("fusions.")]:
        del sys.modules[name]
    import fusions  # noqa: F401
    import fusions.load_llm  # noqa: F401
    import fusions.TTF_RecAvg  # noqa: F401
    import fusions.TTF_T2V_XAttn  # noqa: F401
    import fusions.MMF_GR_Add  # noqa: F401
    import fusions.MMF_XAttn_Add  # noqa: F401
    import fusions.FusionModel  # noqa: F401
    assert os.path.dirname(os.path.abspath(sys.modules["fusions"].__file__)).startswith(PKG)


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    install()
    sys.argv = [script] + sys.argv[2:]
    os.chdir(os.path.dirname(script))  # the reference uses paths relative to its own root (data/, logs/)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
