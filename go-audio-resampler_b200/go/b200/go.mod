module gar-b200/b200

go 1.26
