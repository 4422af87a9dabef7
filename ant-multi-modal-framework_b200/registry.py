"""The plug-in boundary of AntMMF, as seen by the hot path.

If the real `antmmf` package is importable (a user's AntMMF checkout), its registries are used, so that the B200
encoders are constructed by the unmodified `TextEncoder(config)` / `VisualEncoder(config)` calls in
prj/base_vtp/roi_univl/univl/model/univl_video_base.py:24-29. Otherwise (this repository stand-alone, the GPU box) a
behaviour-compatible local ModuleRegistry is used: same `register` / `get` / construction semantics as
antmmf/modules/module_registry.py:9-82 (one dict shared by all registries, keyed by class name; the built module is
exposed as `.module`), and `register_loss` / `register_model` like antmmf/common/registry.py:246-271,415-440.
"""
import inspect

from torch import nn

try:  # pragma: no cover - only with a full AntMMF installation
    from antmmf.common import registry as _antmmf_registry  # type: ignore
    from antmmf.modules.encoders import TextEncoder, VisualEncoder  # type: ignore
    from antmmf.modules.module_registry import ModuleRegistry  # type: ignore

    HAVE_ANTMMF = hasattr(_antmmf_registry, "register_loss")
except Exception:  # noqa: BLE001
    HAVE_ANTMMF = False

if not HAVE_ANTMMF:

    class ModuleRegistry(nn.Module):
        __register_module__ = {}

        @classmethod
        def register(cls, module=None):
            if module is None:
                return cls.register
            if not inspect.isclass(module):
                raise ValueError(f"Only class can be registered, but got {module} with type of `{type(module)}`.")
            cls.__register_module__[module.__name__] = module
            return module

        @classmethod
        def get(cls, module_type):
            if module_type not in cls.__register_module__:
                raise ValueError(f"{module_type} is not registered in {cls.__name__}.")
            return cls.__register_module__[module_type]

        def __init__(self, module_type, *args, **kwargs):
            super().__init__()
            self.module = type(self).get(module_type)(*args, **kwargs)

        def __call__(self, *args, **kwargs):
            return self.module(*args, **kwargs)

    def _cfg_get(config, key, default=None):
        return config.get(key, default) if hasattr(config, "get") else getattr(config, key, default)

    class TextEncoder(ModuleRegistry):
        """antmmf/modules/encoders/text_encoder.py:22-29: built from config.type + config.params."""

        def __init__(self, config, *args, **kwargs):
            params = dict(_cfg_get(config, "params", {}) or {})
            super().__init__(_cfg_get(config, "type"), *args, **params, **kwargs)

    class VisualEncoder(ModuleRegistry):
        """antmmf/modules/encoders/visual_encoder.py:34-50."""

        def __init__(self, config, *args, **kwargs):
            params = dict(_cfg_get(config, "params", {}) or {})
            super().__init__(_cfg_get(config, "type"), *args, **params, **kwargs)

    class _Registry:
        mapping = {"loss_name_mapping": {}, "model_name_mapping": {}}

        @classmethod
        def register_loss(cls, name):
            def wrap(func):
                if not issubclass(func, nn.Module):
                    raise AssertionError("All loss must inherit torch.nn.Module class")
                cls.mapping["loss_name_mapping"][name] = func
                return func

            return wrap

        @classmethod
        def register_model(cls, name):
            def wrap(func):
                cls.mapping["model_name_mapping"][name] = func
                return func

            return wrap

        @classmethod
        def get_loss_class(cls, name):
            return cls.mapping["loss_name_mapping"].get(name)

        @classmethod
        def get_model_class(cls, name):
            return cls.mapping["model_name_mapping"].get(name)

    registry = _Registry
else:  # pragma: no cover
    registry = _antmmf_registry
