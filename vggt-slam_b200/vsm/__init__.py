"""vsm: B200-native semantic voxel mapping + text query for VGGT-SLAM.

Host-side mirror of the reference's ``Submap`` / ``GraphMap`` / ``SemanticVoxel`` /
``SemanticVoxelMap`` API (vggt_slam/submap.py, map.py, semantic_voxel.py) over
the C-ABI library ``csrc/libvsm.so`` (include/vsm.h): hand-written sm_100a CUDA,
no Triton, no CPU fallback.
"""
__version__ = "0.1.0"
