"""Minimal stand-in for the mmcv names the reference imports (mmcv.cnn builders / initialisers, mmcv.runner optimizer
registry) -- same call signatures, torch-only bodies."""
