"""clipb200 -- B200-native hot paths behind CLI-P's `clip` / `faiss` surface.

Put the parent directory (`cli-p_b200/`) on sys.path; it also carries top-level
`clip`, `faiss` drop-in modules so the reference's unchanged `import clip, faiss`
lines resolve to this package.
"""
__version__ = "0.1.0"
