"""B200-native batched IPDDP2 (drop-in for the hot path of mingu6/InteriorPointDDP.jl)."""
