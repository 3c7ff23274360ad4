"""Minimal stand-in for the three timm names the reference imports (models/swin_transformer_v2.py:17)."""
