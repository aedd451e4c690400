from .modeling_clip import CLIPTextModel, CLIPTextModelWithProjection, CLIPVisionModel  # noqa: F401
