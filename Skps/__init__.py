"""Drop-in import surface: `from Skps import FaceAna` (reference: Skps/__init__.py:7)."""
from peppa_pig_face_landmark_b200.core.api.facer import FaceAna

__all__ = ['FaceAna']
