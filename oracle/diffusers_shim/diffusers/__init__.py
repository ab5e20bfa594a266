"""Stand-in for the handful of `diffusers` symbols the reference hot path imports.

TEST INFRASTRUCTURE ONLY (part of oracle/): it lets the reference's own, unmodified
ltx_video/models/transformers/{transformer3d,attention,symmetric_patchifier}.py and
ltx_video/schedulers/rf.py import in a container where diffusers 0.35.1
(/root/reference/requirements.txt:34) is not installable.  Semantics restated from the
published diffusers 0.35.1 behaviour; **parity unpinned** against the
real package because it cannot be installed offline.  Never imported by the product path.
"""
__version__ = "0.35.1-shim"
