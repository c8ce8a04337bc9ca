"""bya_b200: B200-native denoising hot path of Bind-Your-Avatar."""
