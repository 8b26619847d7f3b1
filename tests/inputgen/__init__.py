"""INPUT-GEN: host-side restatements of the reference's input generation (spline fit, Edge.combine, O4H blocking, JSON / CSV
input) -- what is needed to turn the reference's example files into block edges for fixtures and
synthetic workloads.  Test / bench infrastructure, not part of the product package ``turbomesh_b200``."""
