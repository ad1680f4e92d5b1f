"""Stands in for the reference's ``pointnet2`` package (pointnet2/__init__.py is empty there)."""
