"""Einsum expressions (equation, operands, shape) of convolution tensor networks."""
