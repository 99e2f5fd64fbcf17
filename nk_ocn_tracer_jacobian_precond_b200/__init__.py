"""B200-native sparse direct solver for the NK ocean-tracer Jacobian preconditioner."""
