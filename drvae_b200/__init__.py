"""drvae_b200 — B200-native (sm_100a) training hot path of DrVAE / PertVAE / VFAE."""
