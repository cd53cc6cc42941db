"""Test-infrastructure shim package (see nn/conv/__init__.py)."""
