from txt2vid_b200.util import create_object, create_object_file, create_object_json, get_class  # noqa: F401
