"""Place the UNMODIFIED reference files of the hot path (and of the SURVEY 8f "next" rows) under
``baseline/_ref/`` so that they travel to the GPU box with the snapshot.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  ``baseline/_ref/`` is git-ignored (no
reference source ever enters the history) but not gpurun-ignored.  ``/root/reference`` exists
only in the build container; ``__graft_entry__.build()`` calls ``ship()`` there, and
``oracle.ref_loader`` falls back to ``baseline/_ref`` when ``/root/reference`` is absent, so the
``-m gpu`` tests and ``bench.py --impl reference`` can execute the reference itself on the box
(``cpu_baseline.kind == "reference"``).  The reference has no ``setup.py``; this copy IS its
install step (the base contract's ``pip install --target baseline/_ref`` has nothing to build).
"""
import os
import shutil

SRC_ROOT = "/root/reference"
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST_ROOT = os.path.join(REPO_ROOT, "baseline", "_ref")

# Only what the oracle / reference arm executes: torch-only files plus the one native kernel.
FILES = [
    "codebase/utils/functions.py",
    "codebase/utils/monotonic_attention.py",
    "codebase/utils/p_choose_strategy.py",
    "codebase/modules/monotonic_multihead_attention.py",
    "codebase/modules/fixed_pre_decision.py",
    "codebase/models/torch_cif/cif.py",
    "codebase/models/torch_cif/test.py",
    "codebase/models/torch_cif/benchmark.py",
    "codebase/models/cif_transformer.py",
    "codebase/criterion/mma_criterion.py",
    "codebase/criterion/ssnt_loss/ssnt_loss.py",
    "codebase/criterion/ssnt_loss/test.py",
    "codebase/criterion/best_alignment/__init__.py",
    "codebase/criterion/best_alignment/best_alignment.cpp",
    "codebase/criterion/best_alignment/best_alignment.cu",
    "codebase/criterion/best_alignment/LICENSE",
]


def ship(verbose: bool = True) -> bool:
    """Copy FILES from /root/reference to baseline/_ref.  Returns False (and does nothing) when
    the reference tree is not present (i.e. on the GPU box, which uses the shipped copy)."""
    if not os.path.isdir(SRC_ROOT):
        return False
    n = 0
    for rel in FILES:
        src = os.path.join(SRC_ROOT, rel)
        dst = os.path.join(DST_ROOT, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    if verbose:
        print(f"shipped {n} reference files to {DST_ROOT}")
    return True


if __name__ == "__main__":
    ship()
