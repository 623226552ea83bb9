"""vsm: B200-native semantic voxel mapping + text query for VGGT-SLAM.

Host-side mirror of the reference's ``Submap`` / ``GraphMap`` / ``SemanticVoxel`` /
``SemanticVoxelMap`` API (vggt_slam/submap.py, map.py, semantic_voxel.py) over
the C-ABI library ``csrc/libvsm.so`` (include/vsm.h): hand-written sm_100a CUDA,
no Triton, no CPU fallback.

``vsm.synth`` (numpy only) can be imported on its own; everything else needs the
built library and raises ImportError without it.
"""
__version__ = "0.1.0"

_LAZY = {
    "SemanticVoxel": "semantic_voxel", "SemanticVoxelMap": "semantic_voxel", "Submap": "submap", "GraphMap": "map",
    "DeviceVoxelMap": "voxel_map",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(name)
    import importlib

    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
