from .ms_deform_attn_func import MSDeformAttnFunction
from .ms_deform_attn_fused_func import MSDeformAttnFusedFunction, fused_supported
from .layer_epilogue_func import add_layer_norm, linear_relu, zero_masked_rows_

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFusedFunction", "fused_supported",
           "add_layer_norm", "linear_relu", "zero_masked_rows_"]
