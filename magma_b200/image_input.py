"""ImageInput — drop-in for magma/image_input.py: the wrapper `Magma.preprocess_inputs` accepts for images
(magma/magma.py:176-193). Host-side (PIL); not accelerated — the pixels reach the GPU as the [1, 3, R, R] tensor the
model's transform returns."""
from io import BytesIO
from typing import Callable, Union

from PIL import Image as PilImage

_REMOTE_SCHEMES = ("http://", "https://")


def _fetch(url: str, timeout_s: float = 30.0) -> bytes:
    import requests  # imported lazily: local files never need it

    reply = requests.get(url, timeout=timeout_s)
    reply.raise_for_status()
    return reply.content


def load_image(source: Union[str, "PilImage.Image"]) -> "PilImage.Image":
    """A PIL image from an already opened image, a local path, or an http(s) URL."""
    if isinstance(source, PilImage.Image):
        return source
    location = str(source)
    if not location.startswith(_REMOTE_SCHEMES):
        return PilImage.open(location)
    try:
        return PilImage.open(BytesIO(_fetch(location)))
    except Exception as exc:  # the reference's message (image_input.py:19-20), with the cause attached
        raise Exception(f"Could not retrieve image from url:\n{location}") from exc


class ImageInput:
    """magma/image_input.py:6-27: holds `path_or_url` and the decoded `pil_image`; `get_transformed_image` is what
    `Magma.preprocess_inputs` calls with the model's transform (magma/magma.py:186-189). Unlike the reference it also
    takes an opened `PIL.Image.Image`, which is convenient offline."""

    def __init__(self, path_or_url):
        self.path_or_url = path_or_url
        self.pil_image = load_image(path_or_url)

    def get_transformed_image(self, transform_fn: Callable):
        return transform_fn(self.pil_image)
