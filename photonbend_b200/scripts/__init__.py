"""Command line front-end: ``photonbend-b200 make-pano | alter-photo | make-photo`` (or
``python -m photonbend_b200``), with the options of the reference's commands
(photonbend/scripts/commands/*.py, docs/scripts.md).  The commands only do file I/O and
parameter derivation; the remap itself is one fused CUDA kernel on the B200."""
