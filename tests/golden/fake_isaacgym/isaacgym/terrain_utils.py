"""Empty: XBotLCfg uses mesh_type='plane' (reference humanoid_config.py:72)."""
