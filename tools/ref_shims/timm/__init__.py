"""Test-only stand-in for the two `timm` symbols the reference imports at module import time
(networks/utils/convnext_utils.py:29-32, networks/utils/ldm_utils.py:18-21).  It only exists so that
tools/make_golden.py can import the *unmodified* reference in a container without timm; it takes no
part in hot-path arithmetic and is never imported by the product."""
