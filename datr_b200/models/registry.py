"""Model-builder registry: the lookup table main.py:79-85 of the reference uses to find `build_dino`.

API surface kept from the reference's models/registry.py:12-57 because callers rely on it: the module-level
`MODULE_BUILD_FUNCS`, its `registe_with_name(module_name=...)` decorator factory (the reference's spelling),
`register(fn, module_name=None, force=False)`, `get(name)` (None when absent), `name`, `module_dict`, `len()` -- and
the underscore spellings `_module_dict` / `_name`, which main.py:82 reads directly
(`assert args.modelname in MODULE_BUILD_FUNCS._module_dict`)."""
import types
from typing import Callable, Dict, Optional


class Registry:
    """name -> build function.  Only plain functions can be registered; re-registering a name needs force=True."""

    def __init__(self, name: str):
        self.name = name
        self.module_dict: Dict[str, Callable] = {}

    def register(self, module_build_function: Callable, module_name: Optional[str] = None, force: bool = False) -> Callable:
        if not isinstance(module_build_function, types.FunctionType):
            raise TypeError(f"module_build_function must be a function, but got {type(module_build_function)}")
        key = module_build_function.__name__ if module_name is None else module_name
        taken = key in self.module_dict
        if taken and not force:
            raise KeyError(f"{key} is already registered in {self.name}")
        self.module_dict[key] = module_build_function
        return module_build_function

    def registe_with_name(self, module_name: Optional[str] = None, force: bool = False) -> Callable[[Callable], Callable]:
        """Decorator factory: `@MODULE_BUILD_FUNCS.registe_with_name(module_name='dino')`."""
        def decorate(fn: Callable) -> Callable:
            return self.register(fn, module_name=module_name, force=force)
        return decorate

    @property
    def _module_dict(self) -> Dict[str, Callable]:     # main.py:82 / main_teacher.py:82 of the reference
        return self.module_dict

    @property
    def _name(self) -> str:
        return self.name

    def get(self, key: str) -> Optional[Callable]:
        return self.module_dict[key] if key in self.module_dict else None

    def __len__(self) -> int:
        return len(self.module_dict)

    def __repr__(self) -> str:
        return f"{type(self).__name__}(name={self.name}, items={list(self.module_dict)})"


MODULE_BUILD_FUNCS = Registry("model build functions")
