"""Reference-side bindings a maintainer drops into the reference tree."""
