from .build import build
print(build())
