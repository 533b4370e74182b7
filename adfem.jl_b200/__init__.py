"""adfem.jl_b200 — B200-native differentiable FEM assembly (host side)."""
