"""Minimal stand-in for the three `timm.models.layers` symbols the reference imports
(swin_v2_module.py:12, swinfusion_module.py:11).  Harness-side only;
used by make_golden.py to import the unmodified reference in a container without timm."""
