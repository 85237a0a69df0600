"""Puts the B200 path behind the reference's OWN callers (way A).

The north-star contract is that `bokego/mcts.py`, `bokego/gtp.py`, `bin/selfplay.py` and `boke.py` of the reference run
unchanged over this package.  They reach the hot path through exactly two imports -- `bokego.go` and `bokego.nnet`
(/root/reference/bokego/mcts.py:9-11, gtp.py:1-2, boke.py:6-9, bin/selfplay.py:2-3) -- so `install()` registers a package
object named `bokego` whose `go` / `nnet` sub-modules are the mirror modules of this package, and whose search path is the
reference's own `bokego/` directory: every other sub-module (`bokego.mcts`, `bokego.gtp`) is then loaded from the reference's
unmodified files by the normal import machinery.

    import bokego_b200.dropin as dropin
    dropin.install("/path/to/reference")         # directory that holds bokego/mcts.py, bokego/gtp.py
    import bokego.mcts, bokego.gtp               # the reference's files, executing over bokego_b200.go / bokego_b200.nnet

`batch_expansions(MCTS)` additionally injects batched leaf evaluation through the class-level caches the reference already
has (mcts.py:42-44,371-403): whenever the unchanged search expands a node, all of its <= 81 children are
evaluated by ONE encoder launch and ONE policy+value launch and the unchanged `Go_MCTS.features/dist/value` properties hit
the cache afterwards.  Nothing here computes: everything goes to the kernels through bokego_b200.nnet.
"""
import importlib
import importlib.util
import os
import sys
import types

_SAVED = {}


def find_reference(root=None):
    """directory holding the reference's `bokego/` package: `root`, $BOKEGO_REFERENCE, or <repo>/baseline/_ref (where
    tools/install_reference.sh installs it); None when there is none"""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (root, os.environ.get("BOKEGO_REFERENCE"), os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "bokego", "mcts.py")):
            return os.path.abspath(cand)
    return None


def install(reference_root=None):
    """Make `import bokego.go` / `import bokego.nnet` resolve to the B200 mirror and every other `bokego.*` module to the
    reference's own file under `reference_root`.  Returns the package object.  Undo with `uninstall()`."""
    from . import go, nnet
    root = find_reference(reference_root)
    if root is None:
        raise FileNotFoundError("no reference checkout found (pass its directory, set BOKEGO_REFERENCE, or run "
                                "tools/install_reference.sh)")
    for k in [k for k in sys.modules if k == "bokego" or k.startswith("bokego.")]:
        _SAVED.setdefault(k, sys.modules.pop(k))
    pkg = types.ModuleType("bokego")
    pkg.__path__ = [os.path.join(root, "bokego")]
    pkg.__file__ = os.path.join(root, "bokego", "__init__.py")
    pkg.__package__ = "bokego"
    pkg.PKG_PATH = root                                  # bokego/__init__.py:2
    pkg.go, pkg.nnet = go, nnet
    sys.modules["bokego"] = pkg
    sys.modules["bokego.go"] = go
    sys.modules["bokego.nnet"] = nnet
    return pkg


def uninstall():
    """remove the drop-in package again (modules already imported from it keep working)"""
    for k in [k for k in sys.modules if k == "bokego" or k.startswith("bokego.")]:
        del sys.modules[k]
    sys.modules.update(_SAVED)
    _SAVED.clear()


def load_script(path, name):
    """import a reference SCRIPT (bin/selfplay.py, boke.py) as a module without running its `__main__` block"""
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_callers(reference_root=None):
    """(mcts, gtp, selfplay) modules of the reference, loaded unchanged over the mirror; selfplay is None when the checkout
    has no bin/selfplay.py"""
    install(reference_root)
    root = find_reference(reference_root)
    mcts = importlib.import_module("bokego.mcts")
    gtp = importlib.import_module("bokego.gtp")
    sp = os.path.join(root, "bin", "selfplay.py")
    selfplay = load_script(sp, "bokego_ref_selfplay") if os.path.isfile(sp) else None
    return mcts, gtp, selfplay


def batch_expansions(tree_cls, enable=True):
    """Wrap `tree_cls._expand` (mcts.py:185-192) so that the children a node gets are evaluated in one batch and stored in
    the class-level caches (`nnet.prefill_caches`).  The search rule is untouched: the wrapper runs the original `_expand`
    and then only fills caches that `Go_MCTS.features / dist / value` (mcts.py:371-403) would have filled one position at a
    time.  `batch_expansions(cls, False)` restores the original method."""
    from . import nnet
    orig = tree_cls.__dict__.get("_bk_orig_expand")
    if not enable:
        if orig is not None:
            tree_cls._expand = orig
            del tree_cls._bk_orig_expand
        return tree_cls
    if orig is not None:
        return tree_cls
    plain = tree_cls._expand

    def _expand(self, node):
        fresh = node not in self.children
        plain(self, node)
        kids = self.children.get(node)
        if fresh and kids:
            nnet.prefill_caches(tree_cls, list(kids), self.policy_net, self.value_net, device=self.device)

    tree_cls._bk_orig_expand = plain
    tree_cls._expand = _expand
    return tree_cls


def main(argv=None):
    """python -m bokego_b200.dropin <reference script> [args...]: run boke.py / bin/selfplay.py of the reference, unmodified,
    over the B200 mirror (the script sees its own argv)"""
    import runpy
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        raise SystemExit("usage: python -m bokego_b200.dropin <path to boke.py | bin/selfplay.py> [script arguments]")
    script = os.path.abspath(argv[0])
    # boke.py sits in the reference root, bin/selfplay.py one level below it
    root = find_reference(os.path.dirname(script)) or find_reference(os.path.dirname(os.path.dirname(script))) or find_reference()
    install(root)
    sys.argv = [script] + list(argv[1:])
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
