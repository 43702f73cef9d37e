"""ImageInput — drop-in for magma/image_input.py: the wrapper `Magma.preprocess_inputs` accepts for images
(magma/magma.py:176-193). Host-side (PIL); not accelerated — the pixels reach the GPU as the [1, 3, R, R] tensor the
transform returns."""
from io import BytesIO
from typing import Callable

from PIL import Image as PilImage


class ImageInput:
    """An image given as a local path or an http(s) URL (magma/image_input.py:6-27). Also accepts an already opened
    `PIL.Image.Image`, which the reference does not — convenient offline."""

    def __init__(self, path_or_url):
        self.path_or_url = path_or_url
        if isinstance(path_or_url, PilImage.Image):
            self.pil_image = path_or_url
        elif str(path_or_url).startswith(("http://", "https://")):
            try:
                import requests

                self.pil_image = PilImage.open(BytesIO(requests.get(path_or_url, timeout=30).content))
            except Exception as exc:  # same message as the reference (image_input.py:19-20)
                raise Exception(f"Could not retrieve image from url:\n{self.path_or_url}") from exc
        else:
            self.pil_image = PilImage.open(path_or_url)

    def get_transformed_image(self, transform_fn: Callable):
        """Called by Magma.preprocess_inputs with the model's transform (magma/magma.py:186-189)."""
        return transform_fn(self.pil_image)
