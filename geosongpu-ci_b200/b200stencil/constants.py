"""Dimension names (reference: ``ndsl.constants`` as used at dsl_patterns/Do__get_top_of_the_column.py:21)."""
X_DIM, Y_DIM, Z_DIM = "x", "y", "z"
X_INTERFACE_DIM, Y_INTERFACE_DIM, Z_INTERFACE_DIM = "x_interface", "y_interface", "z_interface"
N_HALO_DEFAULT = 3
