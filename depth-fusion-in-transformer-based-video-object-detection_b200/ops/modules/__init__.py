from .ms_deform_attn import MSDeformAttn, project_values

__all__ = ["MSDeformAttn", "project_values"]
