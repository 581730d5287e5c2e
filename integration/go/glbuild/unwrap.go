package glbuild

// Unwrap returns the shader a glbuild wrapper forwards to (nameOverloadShader3D/2D, CachedShader3D/2D,
// overloadBounds3/2: glbuild.go:1095-1330), or nil if s is not a wrapper. The CUDA flattener needs to see through
// the wrappers that ShortenNames3D / OverloadShader3DBounds insert; their unwrap() methods are unexported.
func Unwrap(s Shader) Shader {
	if u, ok := s.(interface{ unwrap() Shader }); ok {
		return u.unwrap()
	}
	return nil
}
