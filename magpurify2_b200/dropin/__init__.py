"""Drop-in installer: makes an installed (pure-Python part of the) reference package use this build.

The reference imports its native functions by module path (magpurify2/tools.py:39-42):

    from magpurify2._codon import get_codon_usage_profile, get_cai, get_codon_index
    from magpurify2._composition import get_gc, get_tnf
    from magpurify2._coverage import get_bam_coverages
    from magpurify2._mmseqs2 import get_mmseqs2

`_composition` and `_coverage` are the two PyO3 extension modules (setup.py:35-36, src/lib.rs:20-21) this
build replaces.  `install(package_dir)` writes two one-line modules of those names into the directory of
the `magpurify2` package; they re-export the functions of magpurify2_b200 (same names, arguments, results),
so `magpurify2 composition` and `magpurify2 coverage` of the reference run unchanged on the GPU kernels.
`_codon` and `_mmseqs2` (codon_usage / taxonomy modules, outside this build) are only written when the
package has no such module, and then as stubs that import fine and fail when CALLED.

    python -m magpurify2_b200.dropin /path/to/site-packages/magpurify2
"""
import os
import sys

_ALIAS = {
    "_composition.py": (
        '"""magpurify2._composition (src/composition.rs:163-168) served by magpurify2_b200 (libmp2.so, sm_100a)."""\n'
        "from magpurify2_b200._composition import get_gc, get_tnf  # noqa: F401\n"),
    "_coverage.py": (
        '"""magpurify2._coverage (src/coverage.rs:208-212) served by magpurify2_b200 (libmp2.so, sm_100a)."""\n'
        "from magpurify2_b200._coverage import get_bam_coverages  # noqa: F401\n"),
}

_STUB = '''"""magpurify2.{name}: not part of the B200 build (the {module} module is outside its scope)."""


def _unavailable(*_a, **_k):
    raise NotImplementedError("magpurify2.{name} is not provided by magpurify2_b200; build the reference's Rust "
                              "crate for the {module} module")


{assignments}
'''

_STUBS = {
    "_codon.py": ("_codon", "codon_usage", ["get_codon_usage_profile", "get_cai", "get_codon_index"]),
    "_mmseqs2.py": ("_mmseqs2", "taxonomy", ["get_mmseqs2"]),
}


def _has_module(package_dir, stem):
    return any(f == stem + ".py" or (f.startswith(stem + ".") and f.endswith((".so", ".pyd")))
               for f in os.listdir(package_dir))


def install(package_dir, stubs=True):
    """Writes the alias modules into `package_dir` (the directory holding magpurify2/__init__.py).
    Existing compiled `_composition` / `_coverage` extensions are shadowed only if they are absent or
    `*.py`; a compiled one is left alone and reported.  Returns the list of files written."""
    package_dir = os.fspath(package_dir)
    if not os.path.isdir(package_dir):
        raise FileNotFoundError(package_dir)
    written = []
    for fname, text in _ALIAS.items():
        stem = fname[:-3]
        compiled = [f for f in os.listdir(package_dir) if f.startswith(stem + ".") and f.endswith((".so", ".pyd"))]
        if compiled:
            raise FileExistsError(f"{package_dir} holds a compiled {stem} ({compiled[0]}): remove it first, a "
                                  "compiled extension wins over a .py module of the same name")
        with open(os.path.join(package_dir, fname), "w") as f:
            f.write(text)
        written.append(fname)
    if stubs:
        for fname, (name, module, funcs) in _STUBS.items():
            if _has_module(package_dir, fname[:-3]):
                continue
            with open(os.path.join(package_dir, fname), "w") as f:
                f.write(_STUB.format(name=name, module=module,
                                     assignments="\n".join(f"{fn} = _unavailable" for fn in funcs)))
            written.append(fname)
    return written


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print(__doc__)
        return 2
    for f in install(argv[0]):
        print("wrote", os.path.join(argv[0], f))
    return 0
