from .utils import calculate_mse, calculate_ssim, images_to_uint8, uint8_to_images  # noqa: F401
