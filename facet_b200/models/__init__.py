"""Model-side mirrors: the CLIP ViT-L/14 image tower + aesthetic head, and the CLIP tagger."""
