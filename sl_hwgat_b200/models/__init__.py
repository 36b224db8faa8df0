"""Drop-in replacements for hwgat/models/{HWGATE,model_params}.py (same file
names, class names, constructor signatures and state_dict keys)."""
