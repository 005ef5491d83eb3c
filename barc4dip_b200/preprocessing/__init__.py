from .normalize import flat_field_correction

__all__ = ["flat_field_correction"]
