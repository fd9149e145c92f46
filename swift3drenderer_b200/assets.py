"""Locates / generates the scene files the harness and bench use."""
from __future__ import annotations

import os

from . import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN_DIR = os.path.join(ROOT, "scenes", "_gen")
SHIPPED = os.path.join(GEN_DIR, "shipped_seed1", "data.bin")


def ensure_shipped_data_bin(seed: int = 1) -> str:
    """data.bin of the reference's demo scene (data-generator/main.swift:375-379) with seeded solid
    orientations.  Textures: the reference's own ppm atlases when /root/reference is present (build
    container; the generated file travels with the repo snapshot), otherwise procedural atlases."""
    path = SHIPPED if seed == 1 else os.path.join(GEN_DIR, f"shipped_seed{seed}", "data.bin")
    if not os.path.exists(path):
        tx, _ = S.default_textures()
        S.write_data_bin(path, S.shipped_scene(seed, tx))
    return path
